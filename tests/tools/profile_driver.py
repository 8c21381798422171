"""Wall time per BO iteration of the full discrete / continuous drivers (hyper-parameter
re-fit every step, as in the option files) and where the host spends it.
Usage: python tests/tools/profile_driver.py [set2d|rand6d|jointbran|conth42] [capital]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

OPTS = {  # the reference's option files, one method (mts family), one trial, short capital
    'set2d': ('d', '--run_id p_set2d\n--trials 1\n--max_capital %d\n--tune_every 1\n--tuning_methods ml\n'
                   '--methods mts\n--functions branin,parabaloid,parabaloid,parabaloid,parabaloid\n'),
    'rand6d': ('d', '--run_id p_rand6d\n--trials 1\n--max_capital %d\n--init_capital 5\n--tune_every 1\n'
                    '--tuning_methods ml-post_sampling\n--hp_samples 5\n--methods mts\n--easy_masses 15\n'
                    '--med_masses 10\n--hard_masses 5\n--massf_dim 6\n'),
    'jointbran': ('d', '--run_id p_jointbran\n--trials 1\n--max_capital %d\n--init_capital 5\n--tune_every 1\n'
                       '--tuning_methods ml\n--methods joint-mts\n--cts_f branin\n--risk_neutral 1\n--num_chops 10\n'),
    'conth42': ('c', '--run_id p_conth42\n--trials 1\n--max_capital %d\n--methods cmts\n--judge_act_size 100\n'
                     '--judge_ctx_size 100\n--function hartmann6\n--act_dim 2\n'),
}


def main():
    import torch
    name = next((a for a in sys.argv[1:] if a in OPTS), 'set2d')
    capital = next((int(a) for a in sys.argv[1:] if a.isdigit()), 40)
    kind, text = OPTS[name]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'o.txt')
        with open(path, 'w') as fh:
            fh.write(text % capital)
        if kind == 'd':
            from ocbo_b200 import ocbo as drv
        else:
            from ocbo_b200 import cts_ocbo as drv
        drv.run(argv=['--options', path])  # warm-up (plans, kernels)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        drv.run(argv=['--options', path])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('=== %s: %d BO iterations in %.3f s = %.2f ms per iteration' % (name, capital, dt, dt / capital * 1e3))
        pr = cProfile.Profile()
        pr.enable()
        drv.run(argv=['--options', path])
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(os.environ.get('SORT', 'cumulative')).print_stats(45)
        print('\n'.join(l[:160] for l in s.getvalue().splitlines()[4:] if l.strip()))


if __name__ == '__main__':
    main()
