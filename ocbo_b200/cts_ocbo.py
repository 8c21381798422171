"""Continuous-context driver: ``python -m ocbo_b200.cts_ocbo --options <file>``.

Host mirror of /root/reference/src/cts_ocbo.py:47-140; reads the reference's option
files unchanged (src/options/conth42.txt, conth22.txt, contbran.txt, *_sethps.txt ...).
Every method of the reference's registry runs (cmts, cmts-pm, pei, revi, rand, ei, ts); a
name outside it raises ``ValueError('Invalid method: ...')`` as at :139 (``--strict_methods 0``
skips it instead).  ``--trials_in_flight`` and rank sharding as in ocbo.py."""
import os
import pickle as pkl
import sys
import time
from argparse import Namespace
from copy import copy

import numpy as np

from . import synth
from .cstrats import cstrats
from .cstrats.cts_opt import cts_opt_args
from .cstrats.baselines import agn_args
from .cstrats.postmax_cts import pm_args
from .cstrats.profile_cts import prof_args
from .options import get_option_specs, load_options
from .util import uniform_draw

ocbo_args = [
    get_option_specs('run_id', False, None, 'Unique ID for the run.'),
    get_option_specs('options', False, None, 'Path to options file.'),
    get_option_specs('trials', False, 3, 'Number of trials to run.'),
    get_option_specs('trials_in_flight', False, 1,
                     'Trials advanced in lockstep on the GPU, their device passes merged (1 = the '
                     'sequential loop of the reference).'),
    get_option_specs('max_capital', False, 50, 'Evaluations per trial after the initial ones.'),
    get_option_specs('init_capital', False, 5, 'Initial evaluations.'),
    get_option_specs('methods', False, None, 'Methods to run, comma separated.'),
    get_option_specs('function', False, None, 'Name of the function to optimise.'),
    get_option_specs('act_dim', False, 1, 'Size of the action domain.'),
    get_option_specs('ctx_constraints', False, None, 'Number of allowed contexts (random).'),
    get_option_specs('hp_tune_samps', False, None, 'Pre-tune HPs on this many samples, then fix.'),
    get_option_specs('write_dir', False, None, 'Directory for per-trial pickles.'),
    get_option_specs('show_fig', False, False, 'Ignored (no plotting).'),
    get_option_specs('write_fig', False, False, 'Ignored (no plotting).'),
    get_option_specs('description', False, None, 'Free-text description.'),
    get_option_specs('set_seed', False, True, 'Fix the global numpy seed.'),
    get_option_specs('strict_methods', False, True, 'Unknown methods raise ValueError like the reference (0: skip them).'),
]
SEED = 100826730
LAST_MERGE_STATS = [None]  # per kind (merged device passes, requests served) of the last in-flight run


def find_function(name):
    for f_info in synth.synth_functions:
        if f_info.name.lower() == str(name).lower():
            return f_info
    raise ValueError('Invalid function: %s' % name)


def assemble_methods(f_info, options):
    if options.methods is None:
        raise ValueError('Methods must be specified.')
    domain = f_info.domain
    ctx_dim = len(domain) - options.act_dim
    slices = None
    if options.ctx_constraints is not None:
        slices = uniform_draw(domain[:ctx_dim], int(options.ctx_constraints))
    methods, eval_set, skipped = [], None, []
    for m_name in options.methods.split(','):
        impl = [s.impl for s in cstrats if s.name.lower() == m_name.lower()]
        if not impl:
            if options.strict_methods:
                raise ValueError('Invalid method: %s' % m_name)
            skipped.append(m_name)
            continue
        methods.append(impl[0](f_info.function, domain, ctx_dim, options, eval_set=eval_set,
                               ctx_constraints=slices))
        eval_set = methods[-1].get_eval_set()
    if skipped:
        sys.stderr.write('ocbo_b200: unknown method names skipped (--strict_methods 0): %s\n' % ','.join(skipped))
    if not methods:
        raise ValueError('No runnable method in: %s' % options.methods)
    return methods


def run_single(methods, f_info, options):
    init_pts = list(uniform_draw(f_info.domain, options.init_capital))
    samps = None if options.hp_tune_samps is None else int(options.hp_tune_samps)
    results = {}
    for method in methods:
        method.set_clean()
        t0 = time.time()
        hist = method.optimize(options.max_capital, init_pts=init_pts, hp_tune_samps=samps)
        hist.time_elapsed = time.time() - t0
        results[method.get_strat_name()] = hist
    return results


def run(argv=None):
    options = load_options(ocbo_args + cts_opt_args + prof_args + pm_args + agn_args, cmd_line=True,
                           argv=argv)
    if options.set_seed:
        np.random.seed(SEED)
    if options.run_id is None:
        raise ValueError('run_id not specified.')
    # One process per GPU under a launcher (torchrun): trial t runs on rank t mod W, no collective
    # inside a trial, ONE gather of the per-trial results; rank 0 writes the pickles.  Every rank
    # assembles the same functions and methods from the same stream (the seed, or rank 0's state).
    from . import dist as _dist
    rank, world_size = _dist.init_from_env()
    if not options.set_seed:
        _dist.share_numpy_stream()  # unseeded: every rank assembles from rank 0's stream
    rand_state = np.random.get_state()
    f_info = find_function(options.function)
    methods = assemble_methods(f_info, options)
    created = None
    if options.write_dir is not None and rank == 0:
        created = os.path.join(options.write_dir, options.run_id)
        if os.path.isdir(created):
            raise ValueError('Run ID already exists: %s' % options.run_id)
        os.makedirs(created)
        pure_f_info = Namespace(**{k: v for k, v in vars(f_info).items() if k != 'function'})
        pure_f_info.function = None  # (:70-72)
        for name, obj in (('run_info.pkl', options), ('rand_state.pkl', rand_state),
                          ('eval_set.pkl', methods[0].get_eval_set()),
                          ('function_info.pkl', pure_f_info)):
            with open(os.path.join(created, name), 'wb') as fh:
                pkl.dump(obj, fh, protocol=2)
    width = max(1, int(options.trials_in_flight or 1))
    LAST_MERGE_STATS[0] = None
    if width > 1 or world_size > 1:
        # trials advance in lockstep and share their device passes (ocbo_b200/inflight.py);
        # their numpy streams are keyed by trial, so the shard a rank runs does not matter
        from . import inflight
        from .gp.gp_interface import B200GPFitter

        def make_trial(t_id):
            own = [copy(m) for m in methods]  # set_clean() gives every copy its own run state
            return lambda: run_single(own, f_info, options)
        mine = _dist.shard_trials(options.trials, rank, world_size)
        local, LAST_MERGE_STATS[0] = inflight.run_waves(
            options.trials, width, SEED, make_trial, [(B200GPFitter, '_fit_counter')], trial_ids=mine)
        all_results = _dist.gather_trials(local, mine, options.trials)
        if all_results is None:
            return None  # not rank 0
    else:
        all_results = []
    for t_id in range(options.trials):
        if t_id == len(all_results):
            all_results.append(run_single(methods, f_info, options))  # the reference's sequential loop
        if created is not None:
            with open(os.path.join(created, 'trial_%d.pkl' % t_id), 'wb') as fh:
                pkl.dump(all_results[t_id], fh, protocol=2)
    return all_results


if __name__ == '__main__':
    run()
