"""Device selection: the "device" is torch's default device, as in the reference
(tedeous/device.py:7-52).  There is no multi-backend dispatch: the fused path runs on CUDA only and
`Solution.evaluate` raises if the default device is not a CUDA device."""
from typing import Any
import torch


def solver_device(device: str):
    """Mirror of tedeous/device.py:7-23 (all tensors created afterwards live on that device)."""
    if device in ('cuda', 'gpu') and torch.cuda.is_available():
        print('CUDA is available and used.')
        return torch.set_default_device('cuda')
    if device in ('cuda', 'gpu'):
        print('CUDA is not available, cpu is used!')
        return torch.set_default_device('cpu')
    print('Default cpu processor is used.')
    return torch.set_default_device('cpu')


def _default_device() -> torch.device:
    return torch.empty(0).device


def check_device(data: Any):
    """Move `data` to the default device (tedeous/device.py:26-46)."""
    device = _default_device()
    if isinstance(data, torch.Tensor):
        return data if data.device == device else data.to(device)
    try:
        return torch.as_tensor(data).to(device)
    except Exception as e:  # same error type as the reference
        raise TypeError(f"Cannot convert data to tensor. Ensure it's a compatible type. Error: {e}")


def device_type() -> str:
    """tedeous/device.py:49-52."""
    return _default_device().type
