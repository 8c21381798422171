"""Discrete / joint driver: ``python -m ocbo_b200.ocbo --options <file>``.

Host mirror of /root/reference/src/ocbo.py:67-144: it reads the reference's option
files UNCHANGED (src/options/set2d.txt, rand6d.txt, jointbran.txt, jointh*.txt ...),
seeds numpy with the reference's seed (:74-75), builds the task functions, runs
``trials`` x methods and (optionally) pickles per-trial results under
``write_dir/run_id`` with the reference's file names (:95-107).  Every method of the
reference's registry runs (MTS / EI families, REVI, the agnostic and random baselines);
a name outside it raises ``ValueError('Invalid method: ...')`` like the reference (:187);
``--strict_methods 0`` reports and skips it instead.  Plotting is out of scope.

Two additions over the reference's sequential loop, both off by default:
``--trials_in_flight R`` advances R trials in lockstep and merges their device passes
(ocbo_b200/inflight.py), and under a launcher (``torchrun -m ocbo_b200.ocbo ...``) the
trials are sharded over the ranks, trial t on rank t mod W, with one gather of the results.
"""
import os
import pickle as pkl
from argparse import Namespace
import sys
from copy import copy, deepcopy

import numpy as np

from .options import get_option_specs, load_options
from .strategies import strategies
from .strategies.multi_opt import multi_opt_args
from . import synth

docbo_args = [
    get_option_specs('run_id', False, None, 'Unique ID for the run.'),
    get_option_specs('options', False, None, 'Path to options file.'),
    get_option_specs('trials', False, 10, 'Number of trials to run.'),
    get_option_specs('trials_in_flight', False, 1,
                     'Trials advanced in lockstep on the GPU, their device passes merged (1 = the '
                     'sequential loop of the reference).'),
    get_option_specs('max_capital', False, 50, 'Evaluations per trial after the initial ones.'),
    get_option_specs('init_capital', False, 5, 'Initial evaluations per task.'),
    get_option_specs('methods', False, None, 'Methods to run, comma separated.'),
    get_option_specs('functions', False, None, 'Named task functions, comma separated.'),
    get_option_specs('cts_f', False, None, '2-D function to chop into 1-D tasks.'),
    get_option_specs('num_chops', False, 5, 'Number of chops of cts_f.'),
    get_option_specs('tune_every', False, 1, 'How often to tune hyper-parameters.'),
    get_option_specs('pretune_for_joint', False, False, 'Pre-tune one joint prior GP.'),
    get_option_specs('tunings', False, None, 'Per-method tuning methods, comma separated.'),
    get_option_specs('write_dir', False, None, 'Directory for per-trial pickles.'),
    get_option_specs('show_fig', False, False, 'Ignored (no plotting).'),
    get_option_specs('write_fig', False, False, 'Ignored (no plotting).'),
    get_option_specs('description', False, None, 'Free-text description.'),
    get_option_specs('max_opt_evals', False, 100, 'Candidate actions per task and step.'),
    get_option_specs('set_seed', False, True, 'Fix the global numpy seed.'),
    get_option_specs('run_ei_after', False, None, 'REVI: switch to round-robin EI after this many steps.'),
    get_option_specs('strict_methods', False, True, 'Unknown methods raise ValueError like the reference (0: skip them).'),
    get_option_specs('chop_h4_2', False, False, 'Hartmann 4 contexts / 2 actions task set.'),
    get_option_specs('chop_h2_2', False, False, 'Hartmann 2 contexts / 2 actions task set.'),
    get_option_specs('chop_h3_1', False, False, 'Hartmann 3 contexts / 1 action task set.'),
    get_option_specs('num_eval_pts', False, 100, 'REVI: judgement actions per task.'),
    get_option_specs('num_candidates', False, 100, 'REVI: candidate actions per task.'),
]
rand_fs_args = [
    get_option_specs('easy_masses', False, 0, 'Easy random mass functions.'),
    get_option_specs('med_masses', False, 0, 'Medium random mass functions.'),
    get_option_specs('hard_masses', False, 0, 'Hard random mass functions.'),
    get_option_specs('global_mass_scale', False, None, 'Common scale of the masses.'),
    get_option_specs('massf_dim', False, 2, 'Dimension of the mass functions.'),
]
SEED = 100826730
LAST_MERGE_STATS = [None]  # per kind (merged device passes, requests served) of the last in-flight run


def assemble(options):
    f_infos = synth.assemble_functions(options.functions)
    scaling = options.global_mass_scale
    scaling = scaling if scaling is None else int(scaling)
    f_infos += synth.random_mass_functions(options.easy_masses, options.med_masses,
                                           options.hard_masses, options.massf_dim, scaling)
    if options.cts_f is not None:
        f_infos += synth.assemble_chopped_1d_functions(options.cts_f, options.num_chops)
    # src/ocbo.py:84-89
    if options.chop_h4_2:
        f_infos += synth.get_chopped_h42()
    if options.chop_h2_2:
        f_infos += synth.get_chopped_h22()
    if options.chop_h3_1:
        f_infos += synth.get_chopped_h31()
    if not f_infos:
        raise ValueError('No functions specified')
    return f_infos


def get_function_gps(f_infos, options):
    """src/ocbo.py:146-159 -- pre-tuned mode: with ``--pretune_for_joint`` ONE empty-data joint
    GP carrying tuned hyper-parameters is shared by every task (and ``tune_every`` drops to 0);
    with ``--tune_every 0`` every task gets its own tuned empty prior (100 d samples); else the
    strategies fit as they go (None)."""
    from .gp.gp_util import get_best_joint_prior, get_best_prior
    if options.pretune_for_joint:
        options.tune_every = 0
        pre_gp = get_best_joint_prior(f_infos)
        return [pre_gp for _ in f_infos]
    if options.tune_every > 0:
        return None
    return [get_best_prior(w.function, w.domain, num_samples=len(w.domain) * 100) for w in f_infos]


def assemble_methods(f_infos, options, gps=None):
    if options.methods is None:
        raise ValueError('Methods must be specified.')
    names = options.methods.split(',')
    tunings = options.tunings.split(',') if options.tunings is not None else None
    if tunings is not None and len(tunings) != len(names):
        raise ValueError('Number of tuning methods must match number of methods.')
    methods, rn_info, skipped = [], None, []
    for m_idx, m_name in enumerate(names):
        impl = [s.impl for s in strategies if s.name.lower() == m_name.lower()]
        if not impl:
            if options.strict_methods:
                raise ValueError('Invalid method: %s' % m_name)
            skipped.append(m_name)
            continue
        if tunings is not None:
            options.tuning_methods = tunings[m_idx]
        methods.append(impl[0](f_infos, options, pre_tuned_gps=gps, rn_info=rn_info))
        rn_info = (methods[-1].rn_grids, methods[-1].rn_bests)
    if skipped:
        sys.stderr.write('ocbo_b200: unknown method names skipped (--strict_methods 0): %s\n' % ','.join(skipped))
    if not methods:
        raise ValueError('No runnable method in: %s' % options.methods)
    if rn_info[1] is not None:  # the random risk-neutral grids give the tasks their maxima (:189-192)
        for f_idx, max_val in enumerate(rn_info[1]):
            f_infos[f_idx].max_val = max_val
    return methods


def run_single(methods, f_infos, options):
    init_pts = []
    for wrap in f_infos:
        low, high = zip(*wrap.domain)
        init_pts.append([np.random.uniform(low, high) for _ in range(options.init_capital)])
    results = {}
    for method in methods:
        method.set_clean()
        hist = method.optimize(options.max_capital, init_pts=init_pts)
        key = method.get_method_name_with_tuning() if options.tunings is not None \
            else method.get_opt_method_name()
        results[key] = hist
    return results


def run(argv=None):
    options = load_options(docbo_args + multi_opt_args + rand_fs_args, cmd_line=True, argv=argv)
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
    f_infos = assemble(options)
    methods = assemble_methods(f_infos, options, get_function_gps(f_infos, options))
    created = None
    if options.write_dir is not None and rank == 0:
        created = os.path.join(options.write_dir, options.run_id)
        if os.path.isdir(created):
            raise ValueError('Run ID already exists: %s' % options.run_id)
        os.makedirs(created)
        # (:229-235) the task Namespaces with ``function = None``: callables do not pickle
        clones = [Namespace(function=None, **deepcopy({k: v for k, v in vars(f).items() if k != 'function'}))
                  for f in f_infos]
        for name, obj in (('run_info.pkl', options), ('rand_state.pkl', rand_state),
                          ('functions.pkl', clones)):
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
            for m in own:  # (the reference's joint strategies SHARE their pre-tuned GP across
                if m.pre_tuned_gps is not None:  # trials, src/strategies/joint_opt.py:52; in
                    m.pre_tuned_gps = deepcopy(m.pre_tuned_gps)  # flight a trial owns its copy)
            return lambda: run_single(own, f_infos, options)
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
            all_results.append(run_single(methods, f_infos, options))  # the reference's sequential loop
        if created is not None:
            with open(os.path.join(created, 'trial_%d.pkl' % t_id), 'wb') as fh:
                pkl.dump(all_results[t_id], fh, protocol=2)
    return all_results


if __name__ == '__main__':
    run()
