"""Full-length runs of the four option-file shapes (capital of the reference's option files: set2d
100, rand6d 1500, jointbran 110, conth42 750), several trials in flight: exercises what the short
tests do not -- training sets growing through several padded sizes, step-plan eviction, ladder
fallbacks late in a run.  Prints iterations/s and the merge statistics.
Usage: python tests/tools/soak.py [name ...] [trials]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CAPITAL = {'set2d': 100, 'rand6d': 1500, 'jointbran': 110, 'conth42': 750}


def main():
    import numpy as np
    import torch
    from tests.tools.profile_driver import OPTS
    names = [a for a in sys.argv[1:] if a in OPTS] or list(OPTS)
    trials = next((int(a) for a in sys.argv[1:] if a.isdigit()), 4)
    for name in names:
        kind, text = OPTS[name]
        if kind == 'd':
            from ocbo_b200 import ocbo as drv
        else:
            from ocbo_b200 import cts_ocbo as drv
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, 'o.txt')
            with open(path, 'w') as fh:
                fh.write(text % CAPITAL[name])
            t0 = time.perf_counter()
            res = drv.run(argv=['--options', path, '--trials', str(trials), '--trials_in_flight', str(trials)])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        key = list(res[0])[0]
        if kind == 'd':
            final = [float(r[key].total_best[-1]) for r in res]
            steps = [len(r[key].choice_timeline) for r in res]
        else:
            final = [float(r[key].score_history[-1].regret) for r in res]
            steps = [len(r[key].query_history) for r in res]
        assert all(s == CAPITAL[name] for s in steps) and np.all(np.isfinite(final))
        print(name, json.dumps({'trials': trials, 'capital': CAPITAL[name], 'wall_s': round(dt, 2),
                                'iterations_per_s': round(trials * CAPITAL[name] / dt, 1),
                                'final': [round(v, 4) for v in final],
                                'merge': {k: v for k, v in drv.LAST_MERGE_STATS[0].items() if k != 'seconds'},
                                'seconds': {k: round(v, 2) for k, v in drv.LAST_MERGE_STATS[0]['seconds'].items()}}),
              flush=True)


if __name__ == '__main__':
    main()
