"""Host-side pieces of bench.py that need no GPU: both arms emit the same `config`, the reference arm never maps the
product library, and the clock sampler reports the timed region (with a stand-in for nvidia-smi)."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    return bench


def test_clock_sampler_keeps_the_rows_of_the_timed_region(tmp_path, monkeypatch):
    fake = tmp_path / 'nvidia-smi'
    # comes up slowly (like the real one on a fresh box), then one row per 50 ms; the SM clock changes at "load"
    fake.write_text('#!/bin/bash\nsleep 0.2\ni=0\nwhile true; do i=$((i+1)); '
                    'if [ $i -le 4 ]; then c=1200; else c=1850; fi; '
                    'echo "0, $c, 1965, 900.0, 0x4, Not Active, Not Active, Not Active, Active"; sleep 0.05; done\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv('PATH', str(tmp_path) + os.pathsep + os.environ['PATH'])
    bench = _bench()
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(1.2)          # the "warm-up": rows at 1200 then 1850 MHz (generous: the stand-in is a bash loop)
    s.mark()
    time.sleep(0.4)
    out = s.stop()
    assert out['window'] == 'timed region' and out['samples'] >= 3
    assert out['sm_mhz'] == 1850.0 and out['sm_max_mhz'] == 1965.0 and out['reasons'] == ['sw_power_cap']
    # a timed region shorter than one sampling period falls back to the last warm-up rows and says so
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(1.0)
    s.t0 = time.perf_counter() + 10.0
    out = s.stop()
    assert out['samples'] >= 1 and out['window'].startswith('warm-up')


def test_reference_arm_config_and_isolation():
    """`--impl reference` prints the product arm's `config` dict and imports neither torch nor the library
    (the arm must be independent of the product)."""
    code = ("import sys, json; sys.argv=['bench.py']; sys.path.insert(0, %r); import bench; "
            "from ocbo_b200.workload import workload_config; "
            "print(json.dumps({'cfg': workload_config(), 'torch': 'torch' in sys.modules, "
            "'lib': 'ocbo_b200._lib' in sys.modules}))" % ROOT)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d['torch'] is False and d['lib'] is False
    assert d['cfg']['workload'].startswith('sweep: n=1024 obs, m=8192 candidates')
