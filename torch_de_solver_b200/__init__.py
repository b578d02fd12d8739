"""tedeous-b200: B200-native residual-loss hot path behind the TEDEouS front-end API."""
from .device import solver_device, check_device, device_type
from .data import Domain, Conditions, Equation

__all__ = ['solver_device', 'check_device', 'device_type', 'Domain', 'Conditions', 'Equation']
