import sys, os, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gc
import torch
from tests.tools.profile_driver import OPTS
from ocbo_b200 import ocbo as drv
MODE = os.environ.get('GC', 'default')
if MODE == 'off':
    gc.disable()
for name in ('set2d', 'rand6d', 'jointbran'):
    kind, text = OPTS[name]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'o.txt')
        open(path, 'w').write(text % 40)
        for R in (1, 16):
            argv = ['--options', path, '--trials', str(R), '--trials_in_flight', str(R)]
            drv.run(argv=argv); torch.cuda.synchronize()
            if MODE == 'freeze':
                gc.collect(); gc.freeze()
            t0 = time.perf_counter(); drv.run(argv=argv); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st = drv.LAST_MERGE_STATS[0]
            print(MODE, name, R, 'wall %.3f s  %.0f it/s' % (dt, R * 40 / dt), json.dumps(st))
