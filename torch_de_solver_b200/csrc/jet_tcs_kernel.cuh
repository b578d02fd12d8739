// Streamed tensor-core path ("tcs"): fused Taylor-jet MLP forward + operator + backward-data kernel for nets with ANY
// number of W x W layers (1 .. 14), warp-specialised, two tiles in flight per CTA.
//
// Same orientation, operand images and 3xTF32 GEMMs as jet_tc_kernel.cuh (D[neuron, (point, channel)] = W . Y, lanes are
// neurons, the tanh-jet rule and its adjoint are thread-local epilogues).  What is different:
//   * the weight gradients of the W x W layers are NOT formed here: every epilogue streams its Y_l / gZ_t rows
//     ([(point, channel) row][neuron], fp32, coalesced) to HBM and wgrad_gemm.cu contracts them afterwards.  Without the
//     dW accumulators TMEM holds only the two D accumulators (256 of 512 columns), without the weight-gradient operand
//     image shared memory has room for the operand images of TWO tiles, and nothing limits the depth of the net.
//   * pre-activation jets saved for the backward sweep live in an L2-resident per-CTA scratch (64 bytes per thread, tile
//     slot and layer, written / read as one coalesced 2 KB run per warp) instead of registers, so no per-tile state
//     crosses an MMA wait in registers and the 16 epilogue warps alternate between the two tile slots: the GEMM of one
//     slot runs behind the epilogue of the other.
//   * warp 16 issues every tcgen05.mma and streams the weight images (two K halves, each reloaded as soon as the last
//     MMA that reads it has retired); the epilogue warps hand over operand images with mbarriers (no CTA-wide barrier
//     between an epilogue and its GEMM).
#pragma once
#include <stdlib.h>
#include "jet_tc_kernel.cuh"
#include "jet_tcs.cuh"

namespace tdb {

constexpr int kTsEpi = 512;                       // epilogue threads (16 warps: 4 lane windows x 4 column parts)
constexpr int kTsThreads = kTsEpi + 96;           // + the MMA / weight-streaming warp + one operator warp per tile slot
constexpr int kSOffW = 0, kSOffAct = 2 * kTcWFloats, kSOffX = kSOffAct + 4 * kTcActFloats,
              kSOffU = kSOffX + 6 * kTcMaxPts * 4, kSOffGu = kSOffU + 2 * kTcMaxOut * kTcCols,      // U, Gu: per slot
              kSOffUP = kSOffGu + 2 * kTcMaxOut * kTcCols, kSOffCg = kSOffUP + 4 * kTcMaxOut * kTcCols,
              kSOffBl = kSOffCg + (kMaxCParams + 3) / 4 * 4, kSOffDbl = kSOffBl + kTcMaxOut,      // Dbl: last-layer bias gradient per slot
              kSOffEnd = kSOffDbl + 2 * kTcMaxOut;
constexpr size_t kTsSmemBytes =
    (size_t)kSOffEnd * 4 + 1024 /*align*/ + 160 /*barriers*/ + kTcMaxTerms * sizeof(tdb200_term) +
    kTcMaxFactors * sizeof(tdb200_factor) + 16 + sizeof(tdb200_segment) + 16 + 32 * 4 + kTcMaxPts * TDB200_MAX_COLS * 8 +
    kTcMaxTerms * 16 + 64;

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// K-steps [s0, s1) of one 3xTF32 layer GEMM: D[:, 0:64] (+)= W_hi Y_hi + W_hi Y_lo + W_lo Y_hi, three N = 64 MMAs into ONE
// accumulator (jet_tc_kernel.cuh issues an N = 128 + an N = 64 MMA and adds two halves in the epilogue: 25 % less tensor
// time, but this kernel is bound by its epilogues and the tensor pipe has slack - a single tcgen05.ld and no adds).
//
// w_tmem != 0: the weights are MMA operands out of TENSOR memory.  Each K-step of the hi / lo image (128 rows x 256 bits,
// the K-major SWIZZLE_128B descriptor of the SS form - profiles/microbench/cp_probe.cu) is copied with tcgen05.cp to
// columns w_tmem + 8 s / w_tmem + 104 + 8 s; tcgen05.cp and tcgen05.mma execute in issue order, so the copies of the NEXT
// layer are issued right behind the last slot's MMAs of the current one and run while the tensor pipe would wait for the
// epilogues; the MMAs then read only B (2 KB) from shared memory (event trace: 42 cycles per MMA against 60-85 for
// shared-memory A operands next to the operand-image stores of the 16 epilogue warps).
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(taddr), "l"(sdesc) : "memory");
}
// One k-block (<= 4 K-steps, compile-time descriptor offsets: the issuing thread is a single instruction stream, and with
// run-time offsets its ~25 dependent uniform-datapath instructions per K-step - not the tensor pipe - set the pace: the
// event trace showed 110 cycles per MMA against 50 for the unrolled form).  awh / awl / bh / bl are the descriptors of the
// k-block's first K-step.
template <int NS>
__device__ __forceinline__ void issue_kblock_steps(uint32_t d_tmem, uint64_t awh, uint64_t awl, uint64_t bh, uint64_t bl,
                                                   bool first_kb, uint32_t wt) {
  constexpr uint32_t idesc64 = umma_idesc(128, kTcCols, 0, 1);
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const uint64_t ao = (uint64_t)(j * 32) >> 4, bo = (uint64_t)(j * 1024) >> 4;
    const uint32_t acc0 = (first_kb && j == 0) ? 0u : 1u;
    if (wt) {
      const uint32_t th = wt + (uint32_t)j * 8, tl = th + kTcWRows;
      umma_tf32_ts(d_tmem, th, bh + bo, idesc64, acc0);
      umma_tf32_ts(d_tmem, th, bl + bo, idesc64, 1u);
      umma_tf32_ts(d_tmem, tl, bh + bo, idesc64, 1u);
    } else {
      umma_tf32(d_tmem, awh + ao, bh + bo, idesc64, acc0);
      umma_tf32(d_tmem, awh + ao, bl + bo, idesc64, 1u);
      umma_tf32(d_tmem, awl + ao, bh + bo, idesc64, 1u);
    }
  }
}
__device__ __forceinline__ void issue_gemm_kblock(uint32_t d_tmem, const float* w_hi, const float* w_lo, const float* b_hi,
                                                  int kb, int nsteps, bool leader, uint32_t w_tmem) {
  const uint64_t awh = umma_desc(smem_u32(w_hi) + (uint32_t)kb * kTcWBlock * 4, 16, 1024),
                 awl = umma_desc(smem_u32(w_lo) + (uint32_t)kb * kTcWBlock * 4, 16, 1024);
  const uint64_t bh = umma_desc(smem_u32(b_hi) + (uint32_t)kb * 4096, kTcActBlock * 4, 512, 1),
                 bl = umma_desc(smem_u32(b_hi + kTcActFloats) + (uint32_t)kb * 4096, kTcActBlock * 4, 512, 1);
  const uint32_t wt = w_tmem ? w_tmem + (uint32_t)kb * 32 : 0u;
  if (leader) {
    if (nsteps >= 4) issue_kblock_steps<4>(d_tmem, awh, awl, bh, bl, kb == 0, wt);
    else if (nsteps == 3) issue_kblock_steps<3>(d_tmem, awh, awl, bh, bl, kb == 0, wt);
    else if (nsteps == 2) issue_kblock_steps<2>(d_tmem, awh, awl, bh, bl, kb == 0, wt);
    else if (nsteps == 1) issue_kblock_steps<1>(d_tmem, awh, awl, bh, bl, kb == 0, wt);
  }
}
// k-block kb (nsteps K-steps) of the weight image pair in shared memory -> tensor memory columns w_tmem + 32 kb ...
// (hi) and + 104 (lo)
__device__ __forceinline__ void issue_weight_copy(uint32_t w_tmem, const float* w_hi, const float* w_lo, int kb, int nsteps,
                                                  bool leader) {
  const uint64_t awh = umma_desc(smem_u32(w_hi) + (uint32_t)kb * kTcWBlock * 4, 16, 1024),
                 awl = umma_desc(smem_u32(w_lo) + (uint32_t)kb * kTcWBlock * 4, 16, 1024);
  if (leader) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nsteps) {
        const uint32_t th = w_tmem + (uint32_t)(kb * 4 + j) * 8;
        tmem_cp_128x256b(th, awh + ((uint64_t)(j * 32) >> 4));
        tmem_cp_128x256b(th + kTcWRows, awl + ((uint64_t)(j * 32) >> 4));
      }
  }
}
// one k-block (32 K values = 4 K-steps; hi and lo image) of a weight image pair -> shared memory
__device__ __forceinline__ void bulk_load_kblock(float* dst, const float* src, int kb, uint64_t* bar) {
  constexpr uint32_t kBytes = kTcWBlock * 4, kChunk = 6656;           // 13312 = 2 x 6656
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(2 * kBytes) : "memory");
#pragma unroll
  for (int img = 0; img < 2; ++img) {
    const uint32_t off = (uint32_t)img * kTcWFloats * 4 + (uint32_t)kb * kBytes;
    for (uint32_t o = 0; o < kBytes; o += kChunk)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(smem_u32(dst) + off + o), "l"(reinterpret_cast<const char*>(src) + off + o), "r"(kChunk), "r"(b)
                   : "memory");
  }
}

#ifdef TDB_TC_TIMING       // phase timers: build with TDB200_TC_TIMING_BUILD=1, run with TDB200_TC_TIMING=1
#define TSMARK(i) do { if (a.dbg) { const long long tn_ = clock64(); tacc[i] += tn_ - tlast; tlast = tn_; \
    if (tr_on && tr_n < 256) tr_buf[tr_n++] = (tn_ << 8) | (i); } } while (0)
#else
#define TSMARK(i) do { } while (0)
#endif

template <int O0, int O1, int O2, bool TM>
__global__ void __launch_bounds__(kTsThreads, 1) jet_tcs_kernel(const JetArgs a, const TcsArgs x) {
  constexpr int J = 1 + O0 + O1 + O2;
  constexpr int ND = (O0 > 0) + (O1 > 0) + (O2 > 0);
  constexpr int PH = kTcPC / J > kTcMaxPts / kTcParts ? kTcMaxPts / kTcParts : kTcPC / J;   // points per column part
  constexpr int P = kTcParts * PH;             // points per tile
  constexpr int C = PH * J;                    // used columns per part (<= 16)
  constexpr int Q = (C + 3) / 4;               // float4 per thread, tile and layer in the streams
  constexpr int ORD[3] = {O0, O1, O2};
  extern __shared__ uint8_t smem_raw_ts[];
  const uint32_t s0_ = smem_u32(smem_raw_ts);
  float* const sbase = reinterpret_cast<float*>(smem_raw_ts + (((s0_ + 1023u) & ~1023u) - s0_));
  uint64_t* bars;
  uint32_t* tmem_ptr;
  tdb200_term* termS; tdb200_factor* facS; tdb200_segment* segS; float* scaleS; double* lossT; int4* recS; int* fastS;
  {
    uint8_t* q = reinterpret_cast<uint8_t*>(sbase + kSOffEnd);
    bars = reinterpret_cast<uint64_t*>(q); q += 160;
    tmem_ptr = reinterpret_cast<uint32_t*>(q); q += 32;
    termS = reinterpret_cast<tdb200_term*>(q); q += (kTcMaxTerms * sizeof(tdb200_term) + 15) / 16 * 16;
    facS = reinterpret_cast<tdb200_factor*>(q); q += (kTcMaxFactors * sizeof(tdb200_factor) + 15) / 16 * 16;
    segS = reinterpret_cast<tdb200_segment*>(q); q += (sizeof(tdb200_segment) + 15) / 16 * 16;
    scaleS = reinterpret_cast<float*>(q); q += 32 * 4;
    lossT = reinterpret_cast<double*>(q); q += kTcMaxPts * TDB200_MAX_COLS * 8;
    recS = reinterpret_cast<int4*>(q); q += kTcMaxTerms * 16;
    fastS = reinterpret_cast<int*>(q);
  }
  uint64_t* const act_full = bars;             // [2] epilogue warps (16 arrivals) -> MMA warp: operand image of the slot is ready
  uint64_t* const d_full = bars + 2;           // [2] MMA warp (tcgen05.commit) -> epilogue warps: accumulator of the slot is ready
  uint64_t* const w_full = bars + 4;           // [4] bulk copies -> MMA warp: k-block of the weight image has landed
  uint64_t* const w_free = bars + 8;           // [4] MMA warp (tcgen05.commit): every MMA reading the k-block has retired
  uint64_t* const op_req = bars + 12;          // [2] epilogue warps (16 arrivals) -> operator warp: last-layer partial sums are in UP
  uint64_t* const op_done = bars + 14;         // [2] operator warp -> epilogue warps: adjoint seeds Gu of the slot are ready
  uint64_t* const up_free = bars + 16;         // operator warp -> epilogue warps: UP has been consumed
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_mma_warp = warp == kTsEpi / 32, is_op_warp = warp == kTsEpi / 32 + 1 || warp == kTsEpi / 32 + 2;
#ifdef TDB_TC_TIMING
  long long tacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tacc[i] = 0;
  long long tlast = clock64();
  bool tr_on = false;                                   // event trace of iteration 40 of CTA 0 (marks: see TSMARK calls)
  int tr_n = 0;
  long long* const tr_buf = a.dbg ? a.dbg + 16 * gridDim.x + (warp == kTsEpi / 32 ? 256 : 0) : nullptr;
#endif
  const int n = (warp & 3) * 32 + lane;                 // neuron = TMEM lane owned by this thread
  const int part = (warp >> 2) & 3;                     // column part (0..3) of the tile this thread owns
  const int L = a.n_layers, W = a.widths[1], n_out = a.widths[L], d = a.d, NM = L - 2;
  const int ksteps = (W + 7) / 8;
  const bool live = n < W;
  const int col0 = part * kTcPC;
  const int G = gridDim.x;
  const int my_tiles = x.tile0 + (int)blockIdx.x < x.tile1 ? (x.tile1 - x.tile0 - (int)blockIdx.x + G - 1) / G : 0;
  const int iters = (my_tiles + 1) / 2;
  const int n_kinds = a.do_grad ? 2 * NM : NM;
  auto tile_of = [&](int it, int slot) { return x.tile0 + (int)blockIdx.x + (2 * it + slot) * G; };
  // launch tile -> (segment of the launch, tile inside that segment); segment 0 is the shared-memory copy
  auto seg_of = [&](int tile, int& m, int& local) -> const tdb200_segment& {
    m = 0;
    while (m + 1 < x.n_msegs && tile >= x.mseg_tile_begin[m + 1]) ++m;
    local = tile - x.mseg_tile_begin[m];
    return m == 0 ? *segS : a.segs[x.mseg_index[m]];
  };
  float* const my_grad = a.part_grad + (size_t)blockIdx.x * a.n_params_pad;    // one partial row per CTA

  // ---- one-time setup --------------------------------------------------------------------------------
  if (x.zero_partials)
    for (int i = tid; i < a.n_params_pad; i += kTsThreads) a.part_grad[(size_t)blockIdx.x * a.n_params_pad + i] = 0.f;
  for (int i = tid; i < 4 * kTcActFloats; i += kTsThreads) (sbase + kSOffAct)[i] = 0.f;    // pad rows / columns stay zero
  if (tid < kMaxCParams) (sbase + kSOffCg)[tid] = 0.f;
  for (int i = tid; i < 6 * kTcMaxPts * 4; i += kTsThreads) (sbase + kSOffX)[i] = 0.f;     // axes >= d stay zero
  for (int i = tid; i < 2 * kTcMaxOut * kTcCols; i += kTsThreads) (sbase + kSOffGu)[i] = 0.f;
  for (int i = tid; i < min(kTcMaxTerms, a.n_terms); i += kTsThreads) termS[i] = a.terms[i];
  for (int i = tid; i < min(kTcMaxFactors, a.n_factors); i += kTsThreads) facS[i] = a.factors[i];
  for (int i = tid; i < (int)(sizeof(tdb200_segment) / 4); i += kTsThreads)
    reinterpret_cast<uint32_t*>(segS)[i] = reinterpret_cast<const uint32_t*>(a.segs + x.mseg_index[0])[i];
  if (tid < a.n_slots) scaleS[tid] = a.slot_scale[tid];
  for (int i = tid; i < kTcMaxPts * TDB200_MAX_COLS; i += kTsThreads) lossT[i] = 0.0;
  if (tid == 0) {
    mbar_init(act_full, 16); mbar_init(act_full + 1, 16);
    for (int i = 2; i < 12; ++i) mbar_init(bars + i, 1);
    mbar_init(op_req, 16); mbar_init(op_req + 1, 16);
    mbar_init(op_done, 1); mbar_init(op_done + 1, 1); mbar_init(up_free, 1);
    *fastS = 1;
  }
  if (is_mma_warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Everything above is independent of the kernel in front (pack_params_kernel, or the weight-gradient GEMM of the
  // previous chunk, which still reads the stream buffers): with a programmatic launch it has run next to it.
  pdl_wait();
  pdl_launch_dependents();             // the weight-gradient GEMM behind sets itself up on the SMs this grid leaves first
  if (tid < kTcMaxOut) (sbase + kSOffBl)[tid] = tid < n_out ? a.arena[a.b_off[L - 1] + tid] : 0.f;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (tid < min(kTcMaxTerms, a.n_terms) && tid < x.term_end) {
    const tdb200_term tm = termS[tid];
    int off[2] = {0xFFFF, 0xFFFF}, ipw[2] = {0, 0}, nf = 0;
    bool ok = tm.kind == 0 || (tm.idx >= 0 && tm.idx < 0x7fffffffLL);
    for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
      const tdb200_factor fc = facS[fi];
      if (fc.ipow == 0) continue;
      if (fc.ipow < 0 || fc.ipow > 3 || nf == 2) { ok = false; break; }
      off[nf] = fc.var * kTcCols + fc.chan;
      ipw[nf] = fc.ipow;
      ++nf;
    }
    if (ok) recS[tid] = make_int4(tm.kind == 0 ? __float_as_int(tm.coeff) : (int)tm.idx, tm.kind, off[0] | (off[1] << 16),
                                  ipw[0] | (ipw[1] << 8));
    else *fastS = 0;
  }
  __syncthreads();

  // =====================================================================================================
  // MMA / weight-streaming warp
  // =====================================================================================================
  if (is_mma_warp) {
    const bool leader = elect_one();
    float* const wbuf = sbase + kSOffW;
    auto image_of = [&](int kind) -> const float* {     // kinds 0..NM-1: W_1..W_NM; NM..2NM-1: W_NM^T..W_1^T
      return kind < NM ? x.wimg + (size_t)kind * 4 * kTcWFloats
                       : x.wimg + (size_t)(2 * NM - 1 - kind) * 4 * kTcWFloats + 2 * kTcWFloats;
    };
    uint32_t act_ph = 0, wfull_ph = 0, wfree_ph = 0;    // one parity bit per slot / k-block
    const uint32_t w_tmem = x.w_in_tmem ? tmem + (512u - 2u * kTcWRows) : 0u;     // columns 304 .. 511: W hi | W lo
    const int n_kb = (ksteps + 3) / 4;
    if (iters > 0 && leader)
      for (int kb = 0; kb < 4; ++kb) bulk_load_kblock(wbuf, image_of(0), kb, w_full + kb);
    __syncwarp();
    if (w_tmem) {
      // ---- weights in tensor memory.  Invariant at the top of job g (= iteration x layer GEMM): the copies of image g are
      // queued or done; shared memory holds nothing that is still needed.  Per job: slot 0's MMAs | request image g + 1
      // (its predecessor's copies have retired by now) | slot 1's MMAs | copies of image g + 1 behind them.
      const int total = iters * n_kinds;
      auto copy_image = [&]() {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(w_full + kb, (wfull_ph >> kb) & 1); wfull_ph ^= 1u << kb;
          tc_fence_after();
          if (kb < n_kb) issue_weight_copy(w_tmem, wbuf, wbuf + kTcWFloats, kb, ksteps - 4 * kb, leader);
          if (leader) umma_commit(w_free + kb);
          __syncwarp();
        }
      };
      if (total > 0) copy_image();
      int g = 0;
      for (int it = 0; it < iters; ++it) {
        const int nslots = (2 * it + 1 < my_tiles) ? 2 : 1;
#ifdef TDB_TC_TIMING
        tr_on = a.dbg && blockIdx.x == 0 && it == 40 && lane == 0;
#endif
        for (int kind = 0; kind < n_kinds; ++kind, ++g) {
          const int next = kind + 1 < n_kinds ? kind + 1 : 0;
          for (int slot = 0; slot < nslots; ++slot) {
            TSMARK(15);
            mbar_wait(act_full + slot, (act_ph >> slot) & 1); act_ph ^= 1u << slot;
            tc_fence_after();
            TSMARK(12);
            const float* b_hi = sbase + kSOffAct + slot * 2 * kTcActFloats;
            const uint32_t dt = tmem + (uint32_t)slot * kTcCols;
#pragma unroll 1
            for (int kb = 0; kb < n_kb; ++kb)
              issue_gemm_kblock(dt, wbuf, wbuf + kTcWFloats, b_hi, kb, ksteps - 4 * kb, leader, w_tmem);
            if (leader) umma_commit(d_full + slot);
            __syncwarp();
            TSMARK(14);
            if (slot == 0 && g + 1 < total) {
#pragma unroll 1
              for (int kb = 0; kb < 4; ++kb) {
                mbar_wait(w_free + kb, (wfree_ph >> kb) & 1); wfree_ph ^= 1u << kb;
                if (leader) bulk_load_kblock(wbuf, image_of(next), kb, w_full + kb);
              }
              __syncwarp();
              TSMARK(13);
            }
          }
          if (g + 1 < total) copy_image();
        }
      }
    } else
    for (int it = 0; it < iters; ++it) {
      const int nslots = (2 * it + 1 < my_tiles) ? 2 : 1;
#ifdef TDB_TC_TIMING
      tr_on = a.dbg && blockIdx.x == 0 && it == 40 && lane == 0;
#endif
      for (int kind = 0; kind < n_kinds; ++kind) {
        const int next = kind + 1 < n_kinds ? kind + 1 : (it + 1 < iters ? 0 : -1);
        for (int slot = 0; slot < nslots; ++slot) {
          TSMARK(15);
          mbar_wait(act_full + slot, (act_ph >> slot) & 1); act_ph ^= 1u << slot;
          tc_fence_after();
          TSMARK(12);
          const float* b_hi = sbase + kSOffAct + slot * 2 * kTcActFloats;
          const uint32_t dt = tmem + (uint32_t)slot * kTcCols;
          const bool last_slot = slot == nslots - 1;     // the shared-memory k-block is free after the LAST slot has read it
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            if (slot == 0) { mbar_wait(w_full + kb, (wfull_ph >> kb) & 1); wfull_ph ^= 1u << kb; }
            TSMARK(13);
            if (kb < n_kb) issue_gemm_kblock(dt, wbuf, wbuf + kTcWFloats, b_hi, kb, ksteps - 4 * kb, leader, 0u);
            if (kb == 3 && leader) umma_commit(d_full + slot);
            if (last_slot && leader) umma_commit(w_free + kb);
            // progressive reload: k-block kb - 1 of the NEXT image is requested as soon as its last reader has retired
            // (the MMAs of k-block kb are queued behind it, so the wait does not starve the tensor pipe)
            if (last_slot && kb > 0 && next >= 0) {
              mbar_wait(w_free + kb - 1, (wfree_ph >> (kb - 1)) & 1); wfree_ph ^= 1u << (kb - 1);
              if (leader) bulk_load_kblock(wbuf, image_of(next), kb - 1, w_full + kb - 1);
            }
            TSMARK(14);
          }
          __syncwarp();
          if (last_slot && next >= 0) {
            mbar_wait(w_free + 3, (wfree_ph >> 3) & 1); wfree_ph ^= 1u << 3;
            if (leader) bulk_load_kblock(wbuf, image_of(next), 3, w_full + 3);
            __syncwarp();
          }
        }
      }
    }
#ifdef TDB_TC_TIMING
    if (a.dbg && lane == 0)
      for (int i = 12; i < 16; ++i) a.dbg[(size_t)blockIdx.x * 16 + i] = tacc[i];
#endif
  } else if (is_op_warp) {
    // ===================================================================================================
    // operator warps (one per tile slot): last-layer sums -> u, operator terms, residual, loss, adjoint seeds (one lane
    // per point)
    // ===================================================================================================
    const bool fast_op = *fastS != 0;
    const int slot = warp - (kTsEpi / 32 + 1);
    uint32_t req_ph = 0;
    double lacc[TDB200_MAX_COLS];                          // per-lane loss sums (runtime column index: local memory)
#pragma unroll
    for (int c = 0; c < TDB200_MAX_COLS; ++c) lacc[c] = 0.0;
    float dbl_lane[kTcMaxOut];                             // last-layer bias gradient: value-channel seeds of this lane's points
#pragma unroll
    for (int v = 0; v < kTcMaxOut; ++v) dbl_lane[v] = 0.f;
    for (int it = 0; it < iters; ++it) {
      const int nslots = (2 * it + 1 < my_tiles) ? 2 : 1;
      if (slot < nslots) {
        int mseg, ltile;
        const tdb200_segment& sg = seg_of(tile_of(it, slot), mseg, ltile);
        const int ncols = sg.n_cols;
        const float* const row_w = mseg == 0 ? a.row_weight : nullptr;     // causal weights: interior rows only
        const long long g_first = (long long)ltile * P;
        const int p_valid = (int)min((long long)P, sg.n_groups - g_first);
        float* const Us = sbase + kSOffU + slot * kTcMaxOut * kTcCols;
        float* const Gus = sbase + kSOffGu + slot * kTcMaxOut * kTcCols;
        mbar_wait(op_req + slot, req_ph); req_ph ^= 1;
        for (int idx = lane; idx < n_out * kTcCols; idx += 32) {
          const int v = idx / kTcCols, r = idx - v * kTcCols;
          const int jc = r & (kTcPC - 1);
          float sacc = (jc < C && jc % J == 0) ? (sbase + kSOffBl)[v] : 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) sacc += (sbase + kSOffUP)[(w * kTcMaxOut + v) * kTcCols + r];
          Us[idx] = sacc;
          Gus[idx] = 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(up_free);
      if (fast_op && ncols > 1) {
        // Systems (Navier-Stokes: 3 equations, ~25 terms): one lane per (point, residual column) instead of one per point
        // - the event trace showed the 16 epilogue warps waiting ~12 k cycles per tile for a warp in which 8 lanes walked
        // all terms.  Values, residuals, seeds and the per-term gradient factors are computed in parallel; only the
        // additions into the adjoint seeds Gu - the columns of one point share entries - happen in a fixed order
        // (term k of column 0, 1, ...): deterministic, and two shared-memory updates per round.
        auto pw = [](float xx, int i) { const float x2 = xx * xx; return i == 1 ? xx : i == 2 ? x2 : i == 3 ? x2 * xx : 1.f; };
        auto dpw = [](float xx, int i) { return i == 1 ? 1.f : i == 2 ? 2.f * xx : i == 3 ? 3.f * xx * xx : 0.f; };
        const int items = p_valid * ncols;
        int tmax = 0;
        for (int c = 0; c < ncols; ++c) tmax = max(tmax, sg.col_term_end[c] - sg.col_term_begin[c]);
        for (int w0 = 0; w0 < items; w0 += 32) {
          const int w = w0 + lane;
          const bool act = w < items;
          const int p = act ? w / ncols : 0, col = act ? w - p * ncols : 0;
          const int pc = (p / PH) * kTcPC + (p % PH) * J;
          const long long row = g_first + p;
          const float* u = Us + pc;
          float* gu = Gus + pc;
          const int tb = sg.col_term_begin[col], te = sg.col_term_end[col];
          float seed = 0.f;
          if (act) {
            float val = 0.f;
            for (int t = tb; t < te; ++t) {
              const int4 r = recS[t];
              const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
              const int o0 = r.z & 0xFFFF, o1 = (r.z >> 16) & 0xFFFF;
              const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
              val = fmaf(cf * pw(x0, r.w & 255), pw(x1, r.w >> 8), val);
            }
            if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
            const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
            const float res = val - tgt;
            const float rw = row_w ? __ldg(row_w + row) : 1.f;
            lacc[sg.col_slot[col] - x.slot_base] += (double)rw * (double)res * (double)res;
            seed = a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col)
                                : 2.f * scaleS[sg.col_slot[col]] * rw * res;
          }
          if (a.do_grad) {
            for (int k = 0; k < tmax; ++k) {
              int o0 = 0xFFFF, o1 = 0xFFFF;
              float v0 = 0.f, v1 = 0.f;
              if (act && tb + k < te) {
                const int4 r = recS[tb + k];
                const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
                o0 = r.z & 0xFFFF; o1 = (r.z >> 16) & 0xFFFF;
                const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
                const float p0 = pw(x0, r.w & 255), p1 = pw(x1, r.w >> 8), sc = seed * cf;
                v0 = sc * dpw(x0, r.w & 255) * p1;
                v1 = sc * p0 * dpw(x1, r.w >> 8);
                if (r.y == 2) atomicAdd(&(sbase + kSOffCg)[r.x], seed * p0 * p1);
              }
              for (int c = 0; c < ncols; ++c) {
                if (act && col == c) {
                  if (o0 != 0xFFFF) gu[o0] += v0;
                  if (o1 != 0xFFFF) gu[o1] += v1;
                }
                __syncwarp();
              }
            }
          }
          __syncwarp();
        }
      } else if (lane < p_valid && fast_op) {
        const int p = lane;
        const int pc = (p / PH) * kTcPC + (p % PH) * J;
        const long long row = g_first + p;
        const float* u = Us + pc;
        float* gu = Gus + pc;
        auto pw = [](float xx, int i) { const float x2 = xx * xx; return i == 1 ? xx : i == 2 ? x2 : i == 3 ? x2 * xx : 1.f; };
        auto dpw = [](float xx, int i) { return i == 1 ? 1.f : i == 2 ? 2.f * xx : i == 3 ? 3.f * xx * xx : 0.f; };
        for (int col = 0; col < ncols; ++col) {
          const int tb = sg.col_term_begin[col], te = sg.col_term_end[col];
          float val = 0.f;
          for (int t = tb; t < te; ++t) {
            const int4 r = recS[t];
            const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
            const int o0 = r.z & 0xFFFF, o1 = (r.z >> 16) & 0xFFFF;
            const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
            val = fmaf(cf * pw(x0, r.w & 255), pw(x1, r.w >> 8), val);
          }
          if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
          const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
          const float res = val - tgt;
          const float rw = row_w ? __ldg(row_w + row) : 1.f;        // causal-loss weight (no grad)
          lacc[sg.col_slot[col] - x.slot_base] += (double)rw * (double)res * (double)res;
          if (!a.do_grad) continue;
          const float seed = a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col)
                                          : 2.f * scaleS[sg.col_slot[col]] * rw * res;
          for (int t = tb; t < te; ++t) {
            const int4 r = recS[t];
            const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
            const int o0 = r.z & 0xFFFF, o1 = (r.z >> 16) & 0xFFFF;
            const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
            const float p0 = pw(x0, r.w & 255), p1 = pw(x1, r.w >> 8), sc = seed * cf;
            if (o0 != 0xFFFF) gu[o0] += sc * dpw(x0, r.w & 255) * p1;
            if (o1 != 0xFFFF) gu[o1] += sc * p0 * dpw(x1, r.w >> 8);
            if (r.y == 2) atomicAdd(&(sbase + kSOffCg)[r.x], seed * p0 * p1);
          }
        }
      } else if (lane < p_valid) {
        const int p = lane;
        const int pc = (p / PH) * kTcPC + (p % PH) * J;
        const long long row = g_first + p;
        for (int col = 0; col < ncols; ++col) {
          float val = 0.f;
          for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
            const tdb200_term tm = termS[t];
            float prod = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                                 : a.arena[a.n_net_params + tm.idx];
            for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
              const tdb200_factor fc = facS[fi];
              prod *= pow_i(Us[fc.var * kTcCols + pc + fc.chan], fc.ipow, fc.pow);
            }
            val += prod;
          }
          if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
          const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
          const float res = val - tgt;
          const float rw = row_w ? __ldg(row_w + row) : 1.f;
          lacc[sg.col_slot[col] - x.slot_base] += (double)rw * (double)res * (double)res;
          if (!a.do_grad) continue;
          const float seed = a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col)
                                          : 2.f * scaleS[sg.col_slot[col]] * rw * res;
          for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
            const tdb200_term tm = termS[t];
            const float cf = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                                     : a.arena[a.n_net_params + tm.idx];
            float full = 1.f;
            for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
              const tdb200_factor fc = facS[fi];
              const float xx = Us[fc.var * kTcCols + pc + fc.chan];
              float part_ = seed * cf * dpow_i(xx, fc.ipow, fc.pow);
              for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
                if (fj == fi) continue;
                const tdb200_factor fo = facS[fj];
                part_ *= pow_i(Us[fo.var * kTcCols + pc + fo.chan], fo.ipow, fo.pow);
              }
              Gus[fc.var * kTcCols + pc + fc.chan] += part_;
              full *= pow_i(xx, fc.ipow, fc.pow);
            }
            if (tm.kind == 2) atomicAdd(&(sbase + kSOffCg)[tm.idx], seed * full);
          }
        }
      }
        if (a.do_grad && lane < p_valid) {
          const int pc = (lane / PH) * kTcPC + (lane % PH) * J;
#pragma unroll
          for (int v = 0; v < kTcMaxOut; ++v) if (v < n_out) dbl_lane[v] += Gus[v * kTcCols + pc];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(op_done + slot);
      }
    }
#pragma unroll
    for (int v = 0; v < kTcMaxOut; ++v) {
      float t = dbl_lane[v];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (lane == 0) (sbase + kSOffDbl)[slot * kTcMaxOut + v] = t;
    }
    if (slot == 0)                                         // per-lane sums by loss slot (relative to x.slot_base)
      for (int c = 0; c < TDB200_MAX_COLS; ++c) lossT[lane * TDB200_MAX_COLS + c] = lacc[c];
    __syncwarp();
    asm volatile("bar.sync 2, 64;" ::: "memory");         // the two operator warps: slot 0 stores, then slot 1 adds
    if (slot == 1)
      for (int c = 0; c < TDB200_MAX_COLS; ++c) lossT[lane * TDB200_MAX_COLS + c] += lacc[c];
  } else {
    // ===================================================================================================
    // epilogue warps
    // ===================================================================================================
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);    // this warp's lane window
    const int actA = (part >> 1) * kTcActBlock + n * 32 + ((((part & 1) * 2) ^ (n & 3)) << 3);       // columns 0..7
    const int actB = (part >> 1) * kTcActBlock + n * 32 + ((((part & 1) * 2 + 1) ^ (n & 3)) << 3);   // columns 8..15
    float w0[4] = {0.f, 0.f, 0.f, 0.f}, wl[kTcMaxOut];
    float dw0_acc[4] = {0.f, 0.f, 0.f, 0.f}, dw0_dir[3] = {0.f, 0.f, 0.f}, dwl_acc[kTcMaxOut];
    float db_acc[kTcsMaxMma + 1];                        // indexed by a runtime layer: lives in local memory (one
#pragma unroll                                           // access per layer and tile)
    for (int l = 0; l <= kTcsMaxMma; ++l) db_acc[l] = 0.f;
    const float bias0 = live ? a.arena[a.b_off[0] + n] : 0.f;
    if (live)
      for (int ax = 0; ax < d; ++ax) w0[ax] = a.arena[a.w_off[0] + n * d + ax];
#pragma unroll
    for (int v = 0; v < kTcMaxOut; ++v) { wl[v] = (live && v < n_out) ? a.arena[a.w_off[L - 1] + v * W + n] : 0.f; dwl_acc[v] = 0.f; }
    const tdb200_segment& sg = *segS;                    // jet directions: those of the launch's first segment
    float w0d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      w0d[i] = i < ND ? fmaf(w0[0], sg.dir_vec[i][0], fmaf(w0[1], sg.dir_vec[i][1], fmaf(w0[2], sg.dir_vec[i][2], w0[3] * sg.dir_vec[i][3])))
                        : 0.f;     // first-order input of direction i: W0[n, :] . v_i (column of W0 for a pure partial)
    uint32_t d_ph = 0, done_ph = 0, up_ph = 0;            // parity bits (per slot) of d_full, op_done; of up_free

    auto xbuf_of = [&](int slot, int it) { return sbase + kSOffX + (slot * 3 + it % 3) * kTcMaxPts * 4; };   // three deep
    auto load_points = [&](int it) {                     // points of both slots of iteration `it` (asynchronous)
      for (int slot = 0; slot < 2; ++slot) {
        const int tile = tile_of(it, slot);
        if (tile >= x.tile1) break;
        float* dst = xbuf_of(slot, it);
        int mseg, ltile;
        const tdb200_segment& st = seg_of(tile, mseg, ltile);
        const long long gf = (long long)ltile * P;
        const int pv = (int)min((long long)P, st.n_groups - gf);
        const float* const src = a.pts + (size_t)(st.pts_off + gf) * d;
        for (int i = tid; i < P * d; i += kTsEpi) {
          const int p = i / d, ax = i - p * d;
          if (p < pv)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst + p * 4 + ax)),
                         "l"(src + (size_t)p * d + ax) : "memory");
          else
            dst[p * 4 + ax] = 0.f;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // tanh-jet rule of this thread's PH points: z (value channel already biased) -> y
    auto jets_fwd = [&](const float* z, bool first, float* y) {
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        const float av = tanh_fast(z[p * J]);
        const TanhF f(av);
        y[p * J] = av;
        int c = 1;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          float zz[4] = {0.f, 0.f, 0.f, 0.f}, yy[4];
          if (first) zz[0] = w0d[i];
          else {
#pragma unroll
            for (int k = 0; k < ORD[i]; ++k) zz[k] = z[p * J + c + k];
          }
          tanh_jet_fwd(f, zz, ORD[i], yy);
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) y[p * J + c + k] = yy[k];
          c += ORD[i];
        }
      }
#pragma unroll
      for (int j = C; j < 16; ++j) y[j] = 0.f;
    };
    // adjoint: gy -> gz (pre-activation cotangents); returns the sum of the value-channel cotangents (bias gradient)
    auto jets_bwd = [&](const float* z, bool first, const float* gy, float* gz, float* g0_out) -> float {
      float db = 0.f;
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        const TanhF f(tanh_fast(z[p * J]));
        float g0 = gy[p * J] * f.f1;
        int c = 1;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          float zz[4] = {0.f, 0.f, 0.f, 0.f}, gg[4];
          if (first) zz[0] = w0d[i];
          else {
#pragma unroll
            for (int k = 0; k < ORD[i]; ++k) zz[k] = z[p * J + c + k];
          }
          g0 += tanh_jet_bwd(f, zz, gy + p * J + c, ORD[i], gg);
          if (first) dw0_dir[i] += gg[0];
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) gz[p * J + c + k] = gg[k];
          c += ORD[i];
        }
        gz[p * J] = g0;
        if (g0_out) g0_out[p] = g0;
        db += g0;
      }
#pragma unroll
      for (int j = C; j < 16; ++j) gz[j] = 0.f;
      return db;
    };
    auto store_act = [&](int slot, const float* v) {      // used columns (Q float4) -> MN-major operand image (hi / lo) of the slot
      float hi[16], lo[16];                               // (the pad columns of the images stay zero from the set-up)
      split16(v, hi, lo);
      float* const ah = sbase + kSOffAct + slot * 2 * kTcActFloats;
      float* const al = ah + kTcActFloats;
      st4(ah + actA, hi); st4(al + actA, lo);
      if (Q > 1) { st4(ah + actA + 4, hi + 4); st4(al + actA + 4, lo + 4); }
      if (Q > 2) { st4(ah + actB, hi + 8); st4(al + actB, lo + 8); }
      if (Q > 3) { st4(ah + actB + 4, hi + 12); st4(al + actB + 4, lo + 12); }
    };
    // this thread's columns -> stream array of one layer: float4 q of (tile, part) at [((tile * 4 + part) * Q + q) * Wp + n],
    // i.e. every warp store writes 512 contiguous bytes.  Which (point, channel) row a column is does not matter to the
    // contraction over rows in wgrad_gemm_kernel as long as Y and gZ use the same order; pad columns and the columns of
    // points beyond the segment carry gZ = 0.
    auto stream_rows = [&](float* arr, int tile, const float* v) {
      if (n >= x.Wp) return;
      float4* dst = reinterpret_cast<float4*>(arr) + ((size_t)(tile - x.tile0) * 4 + part) * (Q * x.Wp) + n;
#pragma unroll
      for (int q = 0; q < Q; ++q) __stcs(dst + q * x.Wp, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    };
    auto hand_over = [&](int slot) {                      // operand image of the slot is complete (this warp's part)
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(act_full + slot);
    };
    auto wait_d = [&](int slot) {
      mbar_wait(d_full + slot, (d_ph >> slot) & 1);
      d_ph ^= 1u << slot;
      tc_fence_after();
    };
    auto load_d = [&](int slot, float* v) { tmem_ld16(t_lane + (uint32_t)slot * kTcCols + (uint32_t)col0, v); };
    // Saved pre-activation jets of W x W layer l (1 .. NM - 1) for the backward sweep.  TM ((NM - 1) * Q <= 12): TMEM
    // columns 128 .. 511 hold one region of 4 parts x 4 Q columns (only the used columns) per saved layer and slot (a
    // tcgen05.st / ld, no memory latency) - three layers of 64 columns, or the four layers of the Navier-Stokes net at 48;
    // other nets use the L2-resident scratch.  A compile-time choice: both paths in one kernel cost 30 registers.
    auto zsave_tmem = [&](int slot, int l) {
      return t_lane + 2 * kTcCols + (uint32_t)(((l - 1) * 2 + slot) * (16 * Q) + part * (4 * Q));
    };
    auto zsave_tmem_st = [&](uint32_t addr, const float* v) {
      if (Q == 4) { tmem_st16(addr, v); return; }
      tmem_st4(addr, v);
      if (Q > 1) tmem_st4(addr + 4, v + 4);
      if (Q > 2) tmem_st4(addr + 8, v + 8);
    };
    auto zsave_tmem_ld = [&](uint32_t addr, float* v) {
      if (Q == 4) { tmem_ld16(addr, v); return; }
      if (Q >= 2) tmem_ld8(addr, v);
      if (Q == 3) {
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr + 8));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) v[8 + i] = __uint_as_float(r[i]);
      }
      if (Q == 1) {
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
      }
#pragma unroll
      for (int j = 4 * Q; j < 16; ++j) v[j] = 0.f;
    };
    auto zsave_ptr = [&](int slot, int l) {                // saved jets of W x W layer l (1..NM)
      return reinterpret_cast<float4*>(x.zsave + ((((size_t)blockIdx.x * 2 + slot) * NM + (l - 1)) * kTsEpi + tid) * 16);
    };

    // layer 0 (K = d) of the slot's tile of iteration it_: thread-local, then the operand image goes to the MMA warp
    auto fwd0 = [&](int it_, int slot) {
      const float* xs = xbuf_of(slot, it_);
      float z[16], y[16];
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + (part * PH + p) * 4);
        z[p * J] = fmaf(w0[0], x4.x, fmaf(w0[1], x4.y, fmaf(w0[2], x4.z, fmaf(w0[3], x4.w, bias0))));
      }
      jets_fwd(z, true, y);
      if (live) store_act(slot, y);
      if (a.do_grad) stream_rows(x.ys, tile_of(it_, slot), y);
      hand_over(slot);
    };
    // With the backward sweep, layer 0 of iteration it + 1 is computed in the TAIL of iteration it, right after the slot's
    // last accumulator has been read (its GEMM then runs behind the thread-local backward of layer 0 and the other slot's
    // tail instead of in front of an idle epilogue).  Forward-only launches keep it at the head of the iteration.
    const bool pipelined = a.do_grad != 0;
    if (iters > 0) load_points(0);
    for (int it = 0; it < iters; ++it) {
      const int nslots = (2 * it + 1 < my_tiles) ? 2 : 1;
#ifdef TDB_TC_TIMING
      tr_on = a.dbg && blockIdx.x == 0 && it == 40 && tid == 0;
#endif
      TSMARK(11);
      if (it == 0 || !pipelined) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        epi_sync();                                       // this iteration's points are visible
        if (it + 1 < iters) load_points(it + 1);
        TSMARK(0);
        for (int slot = 0; slot < nslots; ++slot) { fwd0(it, slot); TSMARK(1); }
      }
      // ---- W x W layers 1 .. NM - 1: GEMM (MMA warp) + tanh-jet epilogue --------------------------------
      for (int l = 1; l < NM; ++l) {
        const float bl = live ? __ldg(a.arena + a.b_off[l] + n) : 0.f;
        for (int slot = 0; slot < nslots; ++slot) {
          float z[16], y[16];
          wait_d(slot);
          TSMARK(2);
          load_d(slot, z);
#pragma unroll
          for (int p = 0; p < PH; ++p) z[p * J] += bl;
          if (a.do_grad) {
            if constexpr (TM) {
              zsave_tmem_st(zsave_tmem(slot, l), z);
            } else {
              float4* zs = zsave_ptr(slot, l);
#pragma unroll
              for (int q = 0; q < 4; ++q) zs[q] = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
            }
          }
          jets_fwd(z, false, y);
          if (live) store_act(slot, y);
          if (a.do_grad) stream_rows(x.ys + (size_t)l * x.stream_stride, tile_of(it, slot), y);
          if (TM && a.do_grad) tmem_st_wait();
          hand_over(slot);
          TSMARK(3);
        }
      }
      // ---- layer NM epilogue + last layer (partial sums for the operator warp) ------------------------------------
      const float bNM = live ? __ldg(a.arena + a.b_off[NM] + n) : 0.f;
      for (int slot = 0; slot < nslots; ++slot) {
        float z[16], y[16];
        wait_d(slot);
        TSMARK(4);
        load_d(slot, z);
#pragma unroll
        for (int p = 0; p < PH; ++p) z[p * J] += bNM;
        jets_fwd(z, false, y);
        if (it > 0 || slot > 0) { mbar_wait(up_free, up_ph); up_ph ^= 1; }     // the operator warp has read the previous sums
        for (int v = 0; v < n_out; ++v) {
          float t16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) t16[j] = wl[v] * y[j];          // wl is zero in dead lanes
          const float tot = warp_multi_reduce16(t16, lane);
          if ((lane & 1) == 0) (sbase + kSOffUP)[((warp & 3) * kTcMaxOut + v) * kTcCols + col0 + reduce16_col(lane)] = tot;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(op_req + slot);
        TSMARK(5);
      }
      if (!a.do_grad) continue;
      // ---- backward of the last layer and of tanh layer NM (the operator warp has produced the adjoint seeds) ------
      for (int slot = 0; slot < nslots; ++slot) {
        const int tile = tile_of(it, slot);
        const float* const Gus = sbase + kSOffGu + slot * kTcMaxOut * kTcCols;
        float z[16], y[16];
        load_d(slot, z);                                  // layer NM's pre-activations are still in the accumulator
#pragma unroll
        for (int p = 0; p < PH; ++p) z[p * J] += bNM;
        jets_fwd(z, false, y);
        mbar_wait(op_done + slot, (done_ph >> slot) & 1); done_ph ^= 1u << slot;
        TSMARK(6);
        float gy[16], gz[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) gy[j] = 0.f;
        for (int v = 0; v < n_out; ++v) {
          float sacc = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 g4 = *reinterpret_cast<const float4*>(Gus + v * kTcCols + col0 + 4 * q);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sacc = fmaf(g[i], y[4 * q + i], sacc);
              gy[4 * q + i] = fmaf(wl[v], g[i], gy[4 * q + i]);
            }
          }
#pragma unroll
          for (int vv = 0; vv < kTcMaxOut; ++vv) if (vv == v) dwl_acc[vv] += sacc;
        }
        db_acc[NM] += jets_bwd(z, false, gy, gz, nullptr);
        if (!live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) gz[j] = 0.f;
        }
        if (live) store_act(slot, gz);
        stream_rows(x.gs + (size_t)(NM - 1) * x.stream_stride, tile, gz);
        hand_over(slot);
        TSMARK(7);
      }
      // ---- backward sweep over the W x W layers NM - 1 .. 1 ---------------------------------------------------
      for (int t = NM - 1; t >= 1; --t) {
        for (int slot = 0; slot < nslots; ++slot) {
          float z[16], gy[16], gz[16];
          if constexpr (!TM) {
            const float4* zs = zsave_ptr(slot, t);        // issued before the wait: the L2 latency hides behind the GEMM
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 v4 = zs[q]; z[4 * q] = v4.x; z[4 * q + 1] = v4.y; z[4 * q + 2] = v4.z; z[4 * q + 3] = v4.w; }
          }
          wait_d(slot);
          TSMARK(8);
          if constexpr (TM) zsave_tmem_ld(zsave_tmem(slot, t), z);
          load_d(slot, gy);
          db_acc[t] += jets_bwd(z, false, gy, gz, nullptr);
          if (!live) {
#pragma unroll
            for (int j = 0; j < 16; ++j) gz[j] = 0.f;
          }
          if (live) store_act(slot, gz);
          stream_rows(x.gs + (size_t)(t - 1) * x.stream_stride, tile_of(it, slot), gz);
          hand_over(slot);
          TSMARK(9);
        }
      }
      // ---- backward of layer 0 (thread-local) --------------------------------------------------------------
      const bool has_next = it + 1 < iters;
      const int nslots_next = (2 * it + 3 < my_tiles) ? 2 : 1;
      if (has_next) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        epi_sync();                                       // points of iteration it + 1 are visible; buffer (it + 2) % 3 is free
        if (it + 2 < iters) load_points(it + 2);
        TSMARK(0);
      }
      for (int slot = 0; slot < nslots; ++slot) {
        const float* xs = xbuf_of(slot, it);
        float z[16], gy[16], gz[16], g0[PH];
        wait_d(slot);
        TSMARK(10);
        load_d(slot, gy);
        if (has_next && slot < nslots_next) { fwd0(it + 1, slot); TSMARK(1); }
#pragma unroll
        for (int p = 0; p < PH; ++p) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + (part * PH + p) * 4);
          z[p * J] = fmaf(w0[0], x4.x, fmaf(w0[1], x4.y, fmaf(w0[2], x4.z, fmaf(w0[3], x4.w, bias0))));
        }
        db_acc[0] += jets_bwd(z, true, gy, gz, g0);
#pragma unroll
        for (int p = 0; p < PH; ++p) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + (part * PH + p) * 4);
          dw0_acc[0] = fmaf(g0[p], x4.x, dw0_acc[0]); dw0_acc[1] = fmaf(g0[p], x4.y, dw0_acc[1]);
          dw0_acc[2] = fmaf(g0[p], x4.z, dw0_acc[2]); dw0_acc[3] = fmaf(g0[p], x4.w, dw0_acc[3]);
        }
      }
    }

    TSMARK(11);
#ifdef TDB_TC_TIMING
    if (a.dbg && tid == 0)
      for (int i = 0; i < 12; ++i) a.dbg[(size_t)blockIdx.x * 16 + i] = tacc[i];
#endif
    // ---- flush: per-thread accumulators of the four column parts -> staging in the (now free) weight buffer; part 0
    // adds them up in a fixed order into the CTA's partial row ----------------------------------------------------
    if (a.do_grad) {
      float* const stg = sbase + kSOffW;                  // [4 parts][NM + 1 + 4 + kTcMaxOut][128]
      const int rows_stg = NM + 1 + 4 + kTcMaxOut;
      for (int i = 0; i < ND; ++i)
        for (int ax = 0; ax < 4; ++ax) dw0_acc[ax] = fmaf(dw0_dir[i], sg.dir_vec[i][ax], dw0_acc[ax]);
      float* mine = stg + (part * rows_stg) * 128 + n;
      for (int l = 0; l <= NM; ++l) mine[l * 128] = db_acc[l];
#pragma unroll
      for (int ax = 0; ax < 4; ++ax) mine[(NM + 1 + ax) * 128] = dw0_acc[ax];
#pragma unroll
      for (int v = 0; v < kTcMaxOut; ++v) mine[(NM + 5 + v) * 128] = dwl_acc[v];
      epi_sync();
      const bool acc = !x.zero_partials;
      auto put = [&](float* q, int r) {
        float t = 0.f;
#pragma unroll
        for (int qq = 0; qq < kTcParts; ++qq) t += stg[(qq * rows_stg + r) * 128 + n];
        *q = acc ? *q + t : t;
      };
      if (live && part == 0) {
        for (int l = 0; l <= NM; ++l) put(my_grad + a.b_off[l] + n, l);
        for (int ax = 0; ax < d; ++ax) put(my_grad + a.w_off[0] + n * d + ax, NM + 1 + ax);
        for (int v = 0; v < n_out; ++v) put(my_grad + a.w_off[L - 1] + v * W + n, NM + 5 + v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---- per-CTA scalars (accumulated by the operator warp) -----------------------------------------------------
  {
    const bool acc = !x.zero_partials;
    if (a.do_grad && tid < n_out) {
      float* q = my_grad + a.b_off[L - 1] + tid;
      const float t = (sbase + kSOffDbl)[tid] + (sbase + kSOffDbl)[kTcMaxOut + tid];
      *q = acc ? *q + t : t;
    }
    if (a.do_grad && tid < a.n_cparams) {
      float* q = my_grad + a.n_net_params + tid;
      *q = acc ? *q + (sbase + kSOffCg)[tid] : (sbase + kSOffCg)[tid];
    }
    if (tid < a.n_slots) {
      double sacc = 0.0;
      const int rel = tid - x.slot_base;
      if (rel >= 0 && rel < TDB200_MAX_COLS)
        for (int p = 0; p < 32; ++p) sacc += lossT[p * TDB200_MAX_COLS + rel];
      double* q = a.part_loss + (size_t)blockIdx.x * a.n_slots + tid;
      *q = acc ? *q + sacc : sacc;
    }
  }
  if (is_mma_warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
template <int O0, int O1, int O2, bool TM>
static cudaError_t launch_tcs_sig(const JetArgs& a, const TcsArgs& x, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_tcs_kernel<O0, O1, O2, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTsSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(jet_tcs_kernel<O0, O1, O2, TM>, dim3(grid), dim3(kTsThreads), kTsSmemBytes, s, a, x);
}

#define TDB_TCS_DEFINE_GROUP(NAME, SIGS)                                                                        \
  cudaError_t NAME(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s) {      \
    SIGS(TDB_TCS_GROUP_CASE)                                                                                    \
    return cudaErrorInvalidValue;                                                                               \
  }
// Tensor-memory budget (512 columns): two 64-column accumulators | saved jets (TM variant: (NM - 1) layers x 2 slots x
// 16 Q columns) | weights (columns 304 .. 511 when x.w_in_tmem).  Measured (wave / Burgers / Navier-Stokes, 10^6 points):
// weights AND saved jets in tensor memory 6.96 / 5.26 ms (shared-memory weights 7.06 / 5.35); where both do not fit
// ((W x W layers - 1) * Q > 5) the saved jets win: Navier-Stokes 25.4 ms with saved jets in TMEM + shared-memory weights,
// 27.0 ms the other way round.  Nets too deep for either ((NM - 1) * Q > 12) keep their saved jets in the L2 scratch and the
// weights in tensor memory.  TDB200_TCS_W_SMEM=1 / TDB200_TCS_NO_TM=1 force the shared-memory / scratch forms.
inline bool tcs_pick(const JetArgs& a, int q, TcsArgs& x) {
  const int regions = (a.n_layers - 3) * q;
  const bool no_tm = getenv("TDB200_TCS_NO_TM") != nullptr, w_smem = getenv("TDB200_TCS_W_SMEM") != nullptr;
  const bool both = regions * 32 <= 512 - 128 - 2 * kTcWRows;
  if (!no_tm && !w_smem && both) { x.w_in_tmem = 1; return true; }
  if (!no_tm && regions <= 12) { x.w_in_tmem = 0; return true; }
  x.w_in_tmem = w_smem ? 0 : 1;
  return false;
}
#define TDB_TCS_GROUP_CASE(A, B, Cc)                                                         \
  if (o0 == A && o1 == B && o2 == Cc) {                                                      \
    TcsArgs xx = x;                                                                          \
    return tcs_pick(a, (jet_tc_columns_per_part(A, B, Cc) + 3) / 4, xx)                      \
               ? launch_tcs_sig<A, B, Cc, true>(a, xx, grid, s)                              \
               : launch_tcs_sig<A, B, Cc, false>(a, xx, grid, s);                            \
  }

}  // namespace tdb
