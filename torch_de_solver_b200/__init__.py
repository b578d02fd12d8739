"""tedeous-b200: B200-native residual-loss hot path behind the TEDEouS front-end API."""
from .device import solver_device, check_device, device_type
from .data import Domain, Conditions, Equation
from .model import Model
from .models import mat_model, parameter_registr
from .optimizers.optimizer import Optimizer
from .solution import Solution

__all__ = ['solver_device', 'check_device', 'device_type', 'Domain', 'Conditions', 'Equation', 'Model',
           'mat_model', 'parameter_registr', 'Optimizer', 'Solution']
