"""
Extrinsic MOSFET wrapper: Rd/Rs, overlap capacitances, drain/source area and
sidewall junctions, parallel multiplier ``m``.

Class factory with the interface of the reference's devices/mosExt.py
(extrinsic_mos :42, process_params :150-248, set_temp_vars :250-279).  The
junction contributions of eval_cqs (:281-324) are part of the MOS kernels
(csrc/models/mosext.cuh); ``cuda_pack`` appends their records to the
intrinsic model's parameter vector.
"""
import numpy as np

from .. import circuit as cir
from ..globalVars import const
from . import cudadev as ad
from .junction import Junction, EMPTY
from .param_layout import flag_bits


def extrinsic_mos(IMOS):
    F = flag_bits(IMOS.cudaModel)

    class ExtMOS(IMOS):
        """
    Extrinsic Silicon MOSFET
    ------------------------

    Adds to the intrinsic model: drain/source resistors (``rsh``, ``nrd``,
    ``nrs``; internal terminals ``di``/``si``), gate overlap capacitances
    (``cgdo``, ``cgso``, ``cgbo``), drain/source area and sidewall junctions
    and the parallel multiplier ``m``.

    Netlist examples::

        {0}:m1 2 3 4 gnd w=10u l=1u asrc=4e-12 ps=8e=12 model=nch
        .model nch {0} (type=n js=1e-3 cj=2e-4 cjsw=1n)
        """
        devType = 'mos' + IMOS.devType.split('_i')[0]
        extraDoc = IMOS.__doc__
        paramDict = dict(
            IMOS.paramDict.items(),
            m=('Parallel multiplier', '', float, 1.),
            cgdo=('Gate-drain overlap capacitance per meter channel width',
                  'F/m', float, 0.),
            cgso=('Gate-source overlap capacitance per meter channel width',
                  'F/m', float, 0.),
            cgbo=('Gate-bulk overlap capacitance per meter channel length',
                  'F/m', float, 0.),
            rsh=('Drain and source diffusion sheet resistance',
                 'Ohm/square', float, 0.),
            js=('Source drain junction current density', 'A/m^2', float, 0.),
            jssw=('Source drain sidewall junction current density', 'A/m',
                  float, 0.),
            pb=('Built in potential of source drain junction', 'V', float,
                .8),
            mj=('Grading coefficient of source drain junction', '', float,
                .5),
            pbsw=('Built in potential of source, drain junction sidewall',
                  'V', float, .8),
            mjsw=('Grading coefficient of source drain junction sidewall',
                  '', float, .33),
            cj=('Source drain junction capacitance per unit area', 'F/m^2',
                float, 0.),
            cjsw=('Source drain junction sidewall capacitance per unit '
                  'length', 'F/m', float, 0.),
            ad=('Drain area', 'm^2', float, 0.),
            asrc=('Source area', 'm^2', float, 0.),
            pd=('Drain perimeter', 'm', float, 0.),
            ps=('Source perimeter', 'm', float, 0.),
            nrd=('Number of squares in drain', 'squares', float, 1.),
            nrs=('Number of squares in source', 'squares', float, 1.),
            fc=('Coefficient for forward-bias depletion capacitances', ' ',
                float, .5),
            xti=('Junction saturation current temperature exponent', '',
                 float, 3.),
            eg0=('Energy bandgap', 'eV', float, 1.11),
        )

        def process_params(self, thermal=False):
            ad.delete_tape(self)
            self.clean_internal_terms()
            self.__addThermalPorts = True
            di, si = 0, 2
            extraVCCS = []
            if self.rsh != 0.:
                if self.nrd != 0.:
                    di = self.add_internal_term('di', 'V')
                    extraVCCS.append(((0, di), (0, di),
                                      1. / self.rsh / self.nrd))
                if self.nrs != 0.:
                    si = self.add_internal_term('si', 'V')
                    extraVCCS.append(((2, si), (2, si),
                                      1. / self.rsh / self.nrs))
            extraVCQS = []
            if self.cgdo != 0.:
                extraVCQS.append(((1, di), (1, di), self.cgdo * self.w))
            if self.cgso != 0.:
                extraVCQS.append(((1, si), (1, si), self.cgso * self.w))
            if self.cgbo != 0.:
                extraVCQS.append(((1, 3), (1, 3), self.cgbo * self.l))
            self.linearVCCS = extraVCCS
            self.linearVCQS = extraVCQS
            if extraVCCS:
                self.csOutPorts = [(di, si), (di, 3), (si, 3)]
                self.controlPorts = [(di, 3), (1, 3), (si, 3)]
                self.qsOutPorts = [(di, 3), (1, 3), (si, 3)]
            self._Tnabs = const.T0 + self.tnom
            self._egapn = self.eg0 - .000702 * (self._Tnabs**2) \
                / (self._Tnabs + 1108.)

            def junction(extent, isat, cj0, vj, m):
                if extent == 0.:
                    return None
                j = Junction()
                j.process_params(isat=isat * extent, cj0=cj0 * extent,
                                 vj=vj, m=m, n=1., fc=self.fc, xti=self.xti,
                                 eg0=self.eg0, Tnomabs=self._Tnabs)
                return j

            self.dj = junction(self.ad, self.js, self.cj, self.pb, self.mj)
            self.sj = junction(self.asrc, self.js, self.cj, self.pb, self.mj)
            self.djsw = junction(self.pd, self.jssw, self.cjsw, self.pbsw,
                                 self.mjsw)
            self.sjsw = junction(self.ps, self.jssw, self.cjsw, self.pbsw,
                                 self.mjsw)
            # intrinsic processing last: it ends with set_temp_vars()
            IMOS.process_params(self, thermal)

        def set_temp_vars(self, temp):
            ad.delete_tape(self)
            IMOS.set_temp_vars(self, temp)
            Tabs = temp + const.T0
            egap_t = self.eg0 - .000702 * (Tabs**2) / (Tabs + 1108.)
            Vt = const.k * Tabs / const.q
            for j in (self.dj, self.djsw, self.sj, self.sjsw):
                if j is not None:
                    j.set_temp_vars(Tabs, self._Tnabs, Vt, self._egapn,
                                    egap_t)

        def cuda_pack(self):
            flags, p = IMOS.cuda_pack(self)
            p = p[:len(p) - 1 - 4 * len(EMPTY)]
            recs = []
            for j, fpres, fcap, cap in (
                    (self.dj, 'AD', 'CJ_D', self.cj),
                    (self.djsw, 'PD', 'CJSW_D', self.cjsw),
                    (self.sj, 'ASRC', 'CJ_S', self.cj),
                    (self.sjsw, 'PS', 'CJSW_S', self.cjsw)):
                if j is None:
                    recs.append(EMPTY)
                    continue
                flags |= F[fpres]
                if cap != 0.:
                    flags |= F[fcap]
                recs.append(j.pack())
            return flags, ad.append_rows(p, [self.m], *recs)

        def cuda_pack_thermal(self):
            """Electro-thermal record: the isothermal record (placeholders at
            ambient for the temperature-dependent entries) followed by the
            raw inputs of the intrinsic model's and of this wrapper's
            set_temp_vars (reference mosExt.py:252-278)."""
            from ..globalVars import glVar
            self.set_temp_vars(glVar.temp)
            flags, base = ExtMOS.cuda_pack(self)
            raw = [j.pack_raw() if j is not None else EMPTY
                   for j in (self.dj, self.djsw, self.sj, self.sjsw)]
            extra = list(IMOS.thermal_raw(self)) \
                + [self._Tnabs, self._egapn, self.eg0]
            return flags, np.concatenate([base, extra] + raw)

        def power(self, vPort, currV):
            return IMOS.power(self, vPort, currV)

    ExtMOS.__doc__ = ExtMOS.__doc__.format(ExtMOS.devType)
    ExtMOS.__name__ = 'ExtMOS_' + IMOS.devType
    return ExtMOS
