"""
Single source of truth for the per-instance parameter vectors the CUDA
device-evaluation kernels consume.

Each model lists, in order, the scalars its kernel reads (post
``process_params``/``set_temp_vars`` values, reference SURVEY App. E).  The
host packers (``cuda_pack`` of each device class) fill vectors in this order;
``python -m cardoon_b200.devices.param_layout`` (run by
``__graft_entry__.build``) writes ``csrc/param_index.h`` with matching enums,
so the two sides cannot drift.
"""

# unified p-n junction record (reference diode.py:20-98, svdiode.py:20-111)
JUNCTION = ['t_is', 'kexp', 't_vj', 'k4', 'k5', 'k6', 'k7', 'fc', 'm',
            'alpha', 'svth', 'k8']
NJ = len(JUNCTION)


def _junc(prefix):
    return [prefix + '_' + n for n in JUNCTION]


MODEL_IDS = dict(diode=0, bjt=1, svbjt=2, bsim3=3, ekv=4, mesfetc=5,
                 memr=6, svdiode=7, diode_t=8, bjt_t=9, svbjt_t=10,
                 mesfetc_t=11, svdiode_t=12, res_t=13, ekv_t=14)

# raw (temperature-independent) junction record of the electro-thermal
# variants: the inputs of Junction.set_temp_vars (reference diode.py:53-72);
# c6 = (1-fc)**(-m-1), c8 = 1 - (1-fc)**k4
JRAW = ['isat', 'n', 'k1', 'k2', 'k3', 'k4', 'vj', 'cj0', 'm', 'fc', 'c6',
        'c8']

DIODE = ['area', 'bv', 'ibv', 'n', 'vt', 'tt'] + _junc('jtn')
DIODE_FLAGS = ['CJ0', 'TT', 'BV', 'QD']

# diode_t: the diode record (its temperature-dependent entries are recomputed
# on the device) followed by the raw inputs of set_temp_vars
DIODE_T = DIODE + ['temp_amb', 'Tnomabs', 'egapn', 'eg0'] \
    + ['jraw_' + n for n in JRAW]

SVDIODE = ['area', 'bv', 'ibv', 'n', 'vt', 'tt', 'gyr'] + _junc('jtn')
SVDIODE_FLAGS = ['CJ0', 'TT', 'BV', 'QD']
SVDIODE_T = SVDIODE + ['temp_amb', 'Tnomabs', 'egapn', 'eg0'] \
    + ['jraw_' + n for n in JRAW]

# res_t (reference resistor.py:107-131): conductance before the temperature
# correction and the coefficients of the correction
RES_T = ['g0', 'tnom', 'tc1', 'tc2', 'temp_amb']

BJT = (['typef', 'area', 'bf_t', 'br_t', 'ikf', 'ikr', 'var', 'vaf', 'rb',
        'rbm', 'ck1', 'ck2', 'tf', 'xtf', 'vtf', 'itf', 'tr', 'xcjc', 'gyr']
       + _junc('jif') + _junc('jir') + _junc('jile') + _junc('jilc')
       + _junc('cbj'))
BJT_FLAGS = ['ISE', 'ISC', 'VAR', 'VAF', 'IKF', 'IKR', 'RB', 'IRB', 'TF',
             'VTF', 'CJE', 'CJC', 'TR', 'QBX', 'QBE', 'QBC', 'SUB', 'CJS']

def _jraw(prefix):
    return [prefix + '_' + n for n in JRAW]


# bjt_t / svbjt_t: the BJT record followed by the raw inputs of set_temp_vars
# (reference bjt.py:453-482, extrinsic :170-180)
BJT_T = BJT + ['temp_amb', 'Tnomabs', 'egapn', 'eg', 'xtb', 'bf', 'br'] \
    + _jraw('rawjif') + _jraw('rawjir') + _jraw('rawjile') \
    + _jraw('rawjilc') + _jraw('rawcbj')

# mosExt wrapper (reference mosExt.py:281-324) appended to both MOS models
MOSEXT = ['m'] + _junc('dj') + _junc('djsw') + _junc('sj') + _junc('sjsw')
MOSEXT_FLAGS = ['AD', 'CJ_D', 'PD', 'CJSW_D', 'ASRC', 'CJ_S', 'PS',
                'CJSW_S']

# BSIM3v3 intrinsic: attributes read by eval_cqs (mosBSIM3v3.py:403-911) with
# the parameter-only sub-expressions folded on the host (see
# mosBSIM3v3.BSIM3.cuda_pack for the definition of each derived name).  The
# order is the order of first use in csrc/models/bsim3.cuh: the evaluation
# kernel streams the record through a shared-memory ring in that order
# (csrc/dev_eval.cu PVRing; the mark() calls in bsim3.cuh name row numbers).
BSIM3 = ['tf', 'vbsc', 'phi', 'sqrtPhi', 'phis3', 'Xdep0', 'V0', 'dvt2',
         'dvt2w', 'factor1', 'c_dvt1', 'dvt0', 'c_dvt1w', 'dvt0w', 't1a',
         'kt1eff', 'kt2', 'ToTnm1', 'eta0', 'etab', 'theta0vb0', 'vth0mk1',
         'k1ox', 'k2ox', 'k3', 'k3b', 'tmp2', 'cox', 'Vt', 'nfepSi', 'cdsc',
         'cdscb', 'cdscd', 'cit', 'cdep0', 'vtcdepcox', 'weff', 'leff',
         'dwg', 'dwb', 'sqrt1nlx', 'prwg', 'prwb', 'rds0', 'xj', 'a0',
         'b0b1', 'agsa0', 'keta', 'tox', 'vsattemp', 'ua', 'uc', 'ub',
         'u0temp', 'a2', 'delta', 'pclm', 'litl', 'thetaRout', 'pdiblb',
         'pvag', 'pscbe2', 'npscbe1litl', 't2sub', 'vfbzb', 'Tox', 'ldeb',
         'epSi', 'noff', 'voffcv', 'avfbzb08', 'acde', 'c1tox', 'c2toxldeb',
         't0dp', 't2dp', 'AbulkCVfactor'] + MOSEXT
BSIM3_FLAGS = MOSEXT_FLAGS + ['PSCBE1', 'ISUB']

# EKV 2.6 intrinsic (mosEKV.py:391-555), same folding convention
EKV = ['tf', 'vt0a', 'deltavRSCE', 'phiT', 'gsqphi', 'gammaa', 'g2o4', 'go2',
       'vt16', 'letaleff', 'weta3', 'weff', 'epscox', 'vt01', 'vt4', 'Vt',
       'vc', 'vdsslog', 'Lambda', 'lamlc', 'lc', 't_ucrit', 'nsleff',
       'lmin2', 'kpanpw', 'bop', 'coxvt', 'e0', 'epSi', 'eta', 'ibn2', 'iba',
       't_ibb', 'ntibblc', 'qscale', 'qox'] + MOSEXT
EKV_FLAGS = MOSEXT_FLAGS + ['SIMPLE']
# mosekv_t: raw inputs of set_temp_vars (reference mosEKV.py:329-357,
# mosExt.py:252-278)
EKV_T = EKV + ['temp_amb', 'vt0', 'tcv', 'Tn', 'avto', 'sga', 'kp', 'bex',
               'akp', 'ucrit', 'ucex', 'phi', 'Tref', 'egTref', 'ibb', 'ibbt',
               'leff', 'ns', 'np', 'cox', 'Cox', 'Tnabs', 'egapn', 'eg0'] \
    + _jraw('rawdj') + _jraw('rawdjsw') + _jraw('rawsj') + _jraw('rawsjsw')

MESFETC = ['area', 'ib0', 'vbd', 'k6', 'Beta', 'vds0', 'a0', 'a1', 'a2',
           'a3', 'gama', 'idsFac', 'Vt0'] + _junc('diogs') + _junc('diogd')
MESFETC_FLAGS = ['CGS0', 'CGD0']
# mesfetc_t: raw inputs of set_temp_vars (reference mesfetc.py:160-189)
MESFETC_T = MESFETC + ['temp_amb', 'Tnomabs', 'egapn', 'eg0', 'tnom', 'nr',
                       'vt0', 'avt0', 'bvt0', 'beta', 'tbet', 'tm', 'tme'] \
    + _jraw('rawgs') + _jraw('rawgd')

# memristor: c, gyr, then the byte-code of M(q) (see memristor.py)
MEMR = ['c', 'gyr']
MEMR_FLAGS = []

LAYOUTS = dict(diode=(DIODE, DIODE_FLAGS), svdiode=(SVDIODE, SVDIODE_FLAGS),
               diode_t=(DIODE_T, DIODE_FLAGS), bjt_t=(BJT_T, BJT_FLAGS),
               svbjt_t=(BJT_T, BJT_FLAGS),
               mesfetc_t=(MESFETC_T, MESFETC_FLAGS),
               svdiode_t=(SVDIODE_T, SVDIODE_FLAGS), res_t=(RES_T, []),
               ekv_t=(EKV_T, EKV_FLAGS),
               bjt=(BJT, BJT_FLAGS), svbjt=(BJT, BJT_FLAGS),
               bsim3=(BSIM3, BSIM3_FLAGS), ekv=(EKV, EKV_FLAGS),
               mesfetc=(MESFETC, MESFETC_FLAGS), memr=(MEMR, MEMR_FLAGS))


def index(model):
    names = LAYOUTS[model][0]
    return {n: i for i, n in enumerate(names)}


def flag_bits(model):
    return {n: (1 << i) for i, n in enumerate(LAYOUTS[model][1])}


def write_header(path):
    lines = ['// Generated by cardoon_b200/devices/param_layout.py -- do not '
             'edit.', '#pragma once', '']
    for model, mid in MODEL_IDS.items():
        lines.append('#define CDN_MODEL_{0} {1}'.format(model.upper(), mid))
    lines.append('#define CDN_NJ {0}'.format(NJ))
    for i, n in enumerate(JUNCTION):
        lines.append('#define J_{0} {1}'.format(n, i))
    for i, n in enumerate(JRAW):
        lines.append('#define JR_{0} {1}'.format(n, i))
    done = set()
    for model, (names, flags) in LAYOUTS.items():
        tag = 'BJT' if model in ('bjt', 'svbjt') else \
            'BJT_T' if model in ('bjt_t', 'svbjt_t') else model.upper()
        if tag in done:
            continue
        done.add(tag)
        lines.append('')
        lines.append('// ---- {0}'.format(tag))
        lines.append('#define NP_{0} {1}'.format(tag, len(names)))
        for i, n in enumerate(names):
            lines.append('#define P_{0}_{1} {2}'.format(tag, n, i))
        for i, n in enumerate(flags):
            lines.append('#define F_{0}_{1} {2}u'.format(tag, n, 1 << i))
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    write_header(os.path.join(here, '..', 'csrc', 'param_index.h'))
