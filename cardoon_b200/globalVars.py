"""
Physical constants (``const``) and the ``.options`` set (``glVar``).

Mirrors the reference's cardoon/globalVars.py:26-60 (same names, units,
defaults) so that netlists and device plugins written for cardoon keep
working.  Values only; nothing here runs on the hot path, the Newton
kernels receive the tolerances as launch arguments.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import math
from .paramset import ParamSet

constDict = dict(
    k=('Boltzmann constant', 'J K^{-1}', float, 1.3806488e-23),
    q=('Elementary charge', 'C', float, 1.602176565e-19),
    epsilon0=('Permittivity of free space', 'F m^{-1}', float,
              8.854187817e-12),
    mu0=('Permeability of free space', 'H m^{-1}', float, math.pi * 4e-7),
    c0=('Speed of light in free space', 'm s^{-1}', float, 2.99792458e8),
    T0=('Zero degree Celsius temperature', 'K', float, 273.15),
    epSi=('Permitivity of silicon', '', float, 104.5e-12),
    epOx=('Permitivity of silicon oxide', '', float, 34.5e-12),
    Np2dB=('Neper to dB conversion constant', 'dB/Np', float,
           8.6858896380650368),
)
const = ParamSet(constDict)
const.set_attributes()

globDict = dict(
    temp=('Ambient temperature', 'C', float, 27.),
    abstol=('Absolute tolerance for nodal variables', 'nodal unit', float,
            1e-7),
    reltol=('Relative tolerance for nodal variables', '', float, 1e-4),
    maxiter=('Maximum number of Newton iterations', '', int, 20),
    maxdelta=('Maximum allowed change in one Newton iteration',
              'nodal unit', float, 20.),
    errfunc=("Enable additional test for error function in Newton's method",
             'bool', bool, False),
    sparse=('Use sparse matrices in analyses if possible', 'bool', bool,
            True),
    gyr=('Default gain in internal gyrators', 'S', float, 1e-3),
)
glVar = ParamSet(globDict)
glVar.set_attributes()
