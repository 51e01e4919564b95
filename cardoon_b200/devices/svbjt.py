"""
State-variable Gummel-Poon BJT (netlist name ``svbjt``), reference
devices/svbjt.py.  The intrinsic class lives in bjt.py next to the plain
model it shares its parameter processing with.
"""
from .bjt import extrinsic_bjt, SVBJTi

Device = extrinsic_bjt(SVBJTi)
