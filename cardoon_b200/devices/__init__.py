"""
Device library: ``devClass`` maps netlist names to device classes.

Mirror of the reference's cardoon/devices/__init__.py (netElemList :39-46,
add_device :54-63): modules listed in ``netElemList`` contribute either a
``Device`` class or a ``devList``; devices with ``makeAutoThermal`` also
register ``<dev>_t``.
"""
import importlib

from . import autoThermal

netElemList = ['linear', 'memristor', 'diode', 'svdiode', 'bjt', 'svbjt',
               'mosEKV',
               'mosBSIM3v3', 'mesfetc', 'tlinpy4', 'tlinps4']

devClass = {}


def add_device(device):
    devClass[device.devType] = device
    # (reference devices/__init__.py:59-63: also for the linear resistor,
    # whose electro-thermal version is nonlinear)
    if device.makeAutoThermal:
        tDevice = autoThermal.thermal_device(device)
        devClass[tDevice.devType] = tDevice


for _modname in netElemList:
    _module = importlib.import_module('.' + _modname, __name__)
    if hasattr(_module, 'devList'):
        for _device in _module.devList:
            add_device(_device)
    else:
        add_device(_module.Device)
