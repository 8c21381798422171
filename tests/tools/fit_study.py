"""Does a bigger hyper-parameter search buy regret?  

Final regret at the option files' FULL capital for the default search (1 round x 64 candidates) against
a 16x larger one (4 rounds x 256: 'ml' zooms on the best candidate, 'post_sampling' draws more), over
several seeds, on the B200 engine (``--engine b200``: ocbo_b200's strategies, needs a GPU) and on the
numpy oracle driven by the REFERENCE's own strategy classes (``--engine oracle``: build container
only, needs /root/reference; rand6d / conth42 at full capital take hours there and are skipped).

    python tests/tools/fit_study.py --engine b200   > gpurun_out/fit_study_b200.json
    python tests/tools/fit_study.py --engine oracle > profiles/r02_fit_study_oracle.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from tests import scenarios as sc  # noqa: E402

SEARCHES = {'64x1': (64, 1), '256x4': (256, 4)}
CONFIGS = {
    'set2d': dict(sc.LONG_DISCRETE['set2d-full']),                       # capital 100
    'jointbran': dict(sc.LONG_DISCRETE['jointbran-full']),               # capital 110
    'rand6d': dict(sc.LONG_DISCRETE['rand6d-full'], capital=1500),       # rand6d.txt:4
    'conth42': dict(sc.LONG_CONTINUOUS['conth42-full-cmts'], capital=750),  # conth42.txt:4
}


def run(engine, names, seeds, searches=None):
    out = {}
    if engine == 'oracle':
        from oracle import gp_numpy, refshim
        refshim.install()
        from strategies.mts import MTS
        from strategies.joint_mts import JointMTS
        from cstrats.profile_cts import ContinuousMultiTaskTS
        from util.misc_util import assemble_functions
        from synth.continuous_2d import assemble_chopped_1d_functions
        from synth.function_generator import get_easy_mass_functions, get_hard_mass_functions, \
            get_med_mass_functions
        from synth.sixd import hartmann6
        classes = {'mts': MTS, 'joint-mts': JointMTS, 'cmts': ContinuousMultiTaskTS}
        counter = gp_numpy.EuclideanGPFitter

        def functions(cfg):
            if 'functions' in cfg:
                return assemble_functions(cfg['functions'])
            if 'cts_f' in cfg:
                return assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
            e, m, h = cfg['masses']
            dom = [[0, 1] for _ in range(cfg['massf_dim'])]
            return get_hard_mass_functions(h, dom, None) + get_med_mass_functions(m, dom, None) + \
                get_easy_mass_functions(e, dom, None)
        gp_engine = 'dragonfly'
    else:
        from ocbo_b200 import synth
        from ocbo_b200.cstrats import cstrats
        from ocbo_b200.gp.gp_interface import B200GPFitter
        from ocbo_b200.strategies import strategies
        classes = {s.name: s.impl for s in strategies}
        classes.update({s.name: s.impl for s in cstrats})
        counter = B200GPFitter
        hartmann6 = synth.hartmann6

        def functions(cfg):
            if 'functions' in cfg:
                return synth.assemble_functions(cfg['functions'])
            if 'cts_f' in cfg:
                return synth.assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
            e, m, h = cfg['masses']
            return synth.random_mass_functions(e, m, h, cfg['massf_dim'])
        gp_engine = 'b200'
    for name in names:
        cfg = CONFIGS[name]
        out[name] = {'capital': cfg['capital'], 'kind': cfg['kind']}
        for tag, (cands, rounds) in SEARCHES.items():
            if searches and tag not in searches:
                continue
            finals, secs = [], []
            for s in range(seeds):
                np.random.seed(sc.SEED + s)
                counter._fit_counter = 0
                if engine == 'oracle':
                    # the reference's strategies build their GP options from the dragonfly option
                    # specs alone (multi_opt.py:67-72): the search size goes in as the specs' default
                    for spec in refshim.EUCLIDEAN_GP_ARGS:
                        if spec['name'] == 'hp_search_candidates':
                            spec['default'] = cands
                        if spec['name'] == 'hp_search_rounds':
                            spec['default'] = rounds
                t0 = time.time()
                if cfg['kind'] == 'cmts':
                    opts = sc.continuous_options(cfg, gp_engine)
                    opts.hp_search_candidates, opts.hp_search_rounds = cands, rounds
                    res = sc.run_continuous(cfg, classes[cfg['kind']], hartmann6,
                                            [[0, 1] for _ in range(6)], opts)
                    finals.append(res['regret'][-1])             # policy regret (cts_opt.py:297-323)
                else:
                    f_infos = functions(cfg)
                    opts = sc.discrete_options(cfg, gp_engine)
                    opts.hp_search_candidates, opts.hp_search_rounds = cands, rounds
                    res = sc.run_discrete(cfg, classes[cfg['kind']], f_infos, opts)
                    best = sum(float(f.max_val) for f in f_infos) if all(
                        getattr(f, 'max_val', None) is not None for f in f_infos) else None
                    # summed simple regret over the tasks if the maxima are known, else -sum of bests
                    finals.append(best - res['total_best'][-1] if best is not None else -res['total_best'][-1])
                secs.append(time.time() - t0)
            out[name][tag] = {'final_regret_per_seed': [float(v) for v in finals],
                              'mean': float(np.mean(finals)), 'std': float(np.std(finals)),
                              'seconds_per_run': float(np.mean(secs)),
                              'lml_evaluations_per_fit': cands + (rounds - 1) * (cands - 1)}
            sys.stderr.write('%s %s %s: %.4f +- %.4f (%.1f s/run)\n' % (
                engine, name, tag, out[name][tag]['mean'], out[name][tag]['std'], out[name][tag]['seconds_per_run']))
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--engine', default='b200', choices=['b200', 'oracle'])
    ap.add_argument('--seeds', type=int, default=5)
    ap.add_argument('--configs', default=None)
    ap.add_argument('--searches', default=None, help='comma separated subset of ' + ','.join(SEARCHES))
    a = ap.parse_args()
    names = a.configs.split(',') if a.configs else (
        ['set2d', 'jointbran'] if a.engine == 'oracle' else list(CONFIGS))
    print(json.dumps({'engine': a.engine, 'seeds': a.seeds, 'searches': SEARCHES, 'results': run(a.engine, names, a.seeds, a.searches.split(',') if a.searches else None)}))
