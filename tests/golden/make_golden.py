"""Generate tests/golden/mts_traces.json by running the REFERENCE's own,
unmodified strategy classes (/root/reference/src, imported where it lies) on
top of the numpy oracle, for the scenarios in tests/scenarios.py.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

from oracle import gp_numpy, refshim  # noqa: E402
from tests import scenarios as sc  # noqa: E402


def long_traces(classes, discrete_functions, hartmann6):
    """tests/golden/mts_long_traces.json: the option files' real task counts and capitals,
    free runs of the reference's own classes (tests/scenarios.py LONG_*)."""
    import time
    out = {}
    for name, cfg in sc.LONG_DISCRETE.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        f_infos = discrete_functions(cfg)
        gp_numpy.TRACE = []
        t0 = time.time()
        r = sc.run_discrete_long(cfg, classes[cfg['kind']], f_infos, sc.discrete_options(cfg, 'dragonfly'))
        trace, gp_numpy.TRACE = gp_numpy.TRACE, None
        per = 1 if cfg['kind'] == 'joint-mts' else len(f_infos)
        r['margins'] = sc.discrete_margins(cfg, len(f_infos), r['choices'], trace)
        r['rungs'] = [[d['jitter_p'] for d in trace[k * per:(k + 1) * per]] for k in range(cfg['capital'])]
        out[name] = r
        nr = sum(1 for st in r['rungs'] for p in st if p is not None)
        print('%s: %d steps in %.1f s, %d of %d draws needed the ladder, %d picks lead by less than their '
              'tolerance' % (name, cfg['capital'], time.time() - t0, nr, per * cfg['capital'],
                             sum(1 for m, t in r['margins'] if m <= t)))
    for name, cfg in sc.LONG_CONTINUOUS.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        gp_numpy.TRACE, gp_numpy.TRACE_EVAL = [], []
        t0 = time.time()
        r = sc.run_continuous_long(cfg, classes[cfg['kind']], hartmann6, [[0, 1] for _ in range(6)],
                                   sc.continuous_options(cfg, 'dragonfly'))
        trace, gp_numpy.TRACE, gp_numpy.TRACE_EVAL = gp_numpy.TRACE, None, None
        P = cfg['num_profiles']
        r['margins'] = sc.continuous_margins(cfg, trace, r['y_init'], r['rewards'])
        r['rungs'] = [[d['jitter_p'] for d in trace[k * P:(k + 1) * P]] for k in range(cfg['capital'])]
        out[name] = r
        nr = sum(1 for st in r['rungs'] for p in st if p is not None)
        print('%s: %d steps in %.1f s, %d of %d draws needed the ladder, final regret %.4f' % (
            name, cfg['capital'], time.time() - t0, nr, P * cfg['capital'], r['regret'][-1]))
    path = os.path.join(ROOT, 'tests', 'golden', 'mts_long_traces.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, separators=(',', ':'))
    print('wrote', path, os.path.getsize(path), 'bytes')


def main():
    refshim.install()
    from strategies.mts import MTS
    from strategies.joint_mts import JointMTS
    from strategies.mei import MEI
    from strategies.joint_mei import JointMEI
    from strategies.agnostic_opt import AgnEI, AgnThompson
    from strategies.joint_agnostic_opt import JointAgnEI, JointAgnThompson
    from strategies.random_strat import JointRandom, RandomOpt
    from cstrats.profile_cts import CMTSPM, ContinuousMultiTaskTS, ProfileEI
    from cstrats.agn_cts import AgnEI as CtsAgnEI, AgnTS as CtsAgnTS
    from cstrats.rand_cts import RandOpt
    if not hasattr(np, 'float'):
        np.float = float  # numpy >= 1.24 removed the alias misc_util.py:192 uses (REVI only)
    from strategies.corr_opt import REVI
    from cstrats.postmax_cts import REVI as CtsREVI
    from util.misc_util import assemble_functions
    from synth.continuous_2d import assemble_chopped_1d_functions
    from synth.function_generator import get_easy_mass_functions, get_hard_mass_functions, \
        get_med_mass_functions
    from synth.sixd import hartmann6

    def discrete_functions(cfg):
        if 'functions' in cfg:
            return assemble_functions(cfg['functions'])
        return assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])

    def functions_of(cfg):
        if 'functions' in cfg:
            return assemble_functions(cfg['functions'])
        if 'cts_f' in cfg:
            return assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
        easy, med, hard = cfg['masses']
        domain = [[0, 1] for _ in range(cfg['massf_dim'])]
        return get_hard_mass_functions(hard, domain, None) + \
            get_med_mass_functions(med, domain, None) + get_easy_mass_functions(easy, domain, None)

    if '--short-only' not in sys.argv:
        long_traces({'mts': MTS, 'joint-mts': JointMTS, 'cmts': ContinuousMultiTaskTS, 'cmts-pm': CMTSPM},
                    functions_of, hartmann6)
    if '--long-only' in sys.argv:
        return
    out = {}
    # baselines: picks only (no tie margins: the CPU host-logic test is their parity gate)
    base_cls = {'agn-thomp': AgnThompson, 'agn-ei': AgnEI, 'Random': RandomOpt, 'ja-thomp': JointAgnThompson,
                'ja-ei': JointAgnEI, 'joint-rand': JointRandom, 'revi': REVI}
    for name, cfg in sc.DISCRETE_BASELINES.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        f_infos = discrete_functions(cfg)
        out[name] = sc.run_discrete(cfg, base_cls[cfg['kind']], f_infos, sc.discrete_options(cfg, 'dragonfly'))
        print(name, out[name]['choice_timeline'])
    for name, cfg in sc.CONTINUOUS_BASELINES.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        cls = {'rand': RandOpt, 'ts': CtsAgnTS, 'ei': CtsAgnEI, 'revi': CtsREVI}[cfg['kind']]
        out[name] = sc.run_continuous(cfg, cls, hartmann6, [[0, 1] for _ in range(6)],
                                      sc.continuous_options(cfg, 'dragonfly'))
        print(name, np.round(out[name]['regret'], 4).tolist())
    for name, cfg in sc.DISCRETE.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        if 'functions' in cfg:
            f_infos = assemble_functions(cfg['functions'])
        elif 'cts_f' in cfg:
            f_infos = assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
        else:
            easy, med, hard = cfg['masses']
            domain = [[0, 1] for _ in range(cfg['massf_dim'])]
            f_infos = get_hard_mass_functions(hard, domain, None) + \
                get_med_mass_functions(med, domain, None) + get_easy_mass_functions(easy, domain, None)
        cls = {'mts': MTS, 'joint-mts': JointMTS, 'mei': MEI, 'joint-mei': JointMEI}[cfg['kind']]
        gp_numpy.TRACE = []
        out[name] = sc.run_discrete(cfg, cls, f_infos, sc.discrete_options(cfg, 'dragonfly'))
        out[name]['margins'] = sc.discrete_margins(cfg, len(f_infos), out[name]['choice_timeline'],
                                                   gp_numpy.TRACE)
        out[name]['jitter_rungs'] = [d['jitter_p'] for d in gp_numpy.TRACE]
        gp_numpy.TRACE = None
        print(name, out[name]['choice_timeline'],
              ['%.1e/%.1e' % m for m in out[name]['margins']])
    for name, cfg in sc.CONTINUOUS.items():
        np.random.seed(sc.SEED)
        gp_numpy.EuclideanGPFitter._fit_counter = 0
        domain = [[0, 1] for _ in range(6)]
        cls = {'cmts': ContinuousMultiTaskTS, 'cmts-pm': CMTSPM, 'pei': ProfileEI}[cfg['kind']]
        gp_numpy.TRACE, gp_numpy.TRACE_EVAL = [], []
        out[name] = sc.run_continuous(cfg, cls, hartmann6, domain,
                                      sc.continuous_options(cfg, 'dragonfly'))
        trace = gp_numpy.TRACE_EVAL if cfg['kind'] == 'pei' else gp_numpy.TRACE
        gp_numpy.TRACE_EVAL = None
        out[name]['margins'] = sc.continuous_margins(cfg, trace, out[name]['y_init'],
                                                     out[name]['rewards'])
        out[name]['jitter_rungs'] = [d['jitter_p'] for d in gp_numpy.TRACE]
        gp_numpy.TRACE = None
        print(name, np.round(out[name]['regret'], 4).tolist(),
              ['%.1e/%.1e' % m for m in out[name]['margins']])
    path = os.path.join(ROOT, 'tests', 'golden', 'mts_traces.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=1)
    print('wrote', path)


if __name__ == '__main__':
    main()
