"""R independent BO trials in flight on one GPU.

The reference runs its trials one after the other (src/ocbo.py:104-107,
src/cts_ocbo.py:73-76) and every trial's decision is a handful of problems of
<= 205 rows (set2d: 5 tasks): one trial alone leaves a B200 idle.  Trials share
no state, so BASELINE.json's north star shards the work "by independent BO
trials"; at the option-file shapes that axis is used INSIDE a GPU: the trials
advance in lockstep and the device passes they ask for at the same point of
their loops -- the MTS / MEI decision over all tasks, the LML pass of a
hyper-parameter fit -- are merged into ONE batched launch sequence (the batched
kernels do not care which trial a problem belongs to; padding is exact, so every
trial gets the bits it would get alone).

Mechanism: each trial runs the UNCHANGED strategy code in its own thread, but
only one thread runs at a time (a baton handed over by the scheduler in the
calling thread), so the host code stays single-threaded as the reference's is.
A trial that reaches a merge point (``merge``) parks its request and hands the
baton back; when no trial is runnable the scheduler executes the most requested
kind for all parked trials at once, in its own thread, and hands the results
out.  Process-global state that belongs to a trial -- the global numpy stream the
reference draws from, the fitter's criterion counter -- is switched at every
hand-over, so a trial consumes ITS stream exactly as it would alone.

The global numpy stream is virtualised rather than copied: ``np.random.uniform``
& co. are module attributes bound to one hidden ``RandomState``; every trial owns
a ``RandomState`` of its own and a hand-over re-points those ~50 module attributes
at the running trial's generator (~5 us; ``get_state`` / ``set_state`` cost ~55 us
EACH, more than the rest of a hand-over).  Code that spells ``np.random.<f>(...)``
-- the reference's strategies and their mirrors do -- is served; a function object
imported BEFORE the run (``from numpy.random import uniform``) would still draw
from the process-wide stream.
"""
import threading
import time
from collections import deque

import numpy as np

_tls = threading.local()

# the module-level functions of numpy's legacy global stream (seed, uniform, normal, ...)
_STREAM_NAMES = [n for n in dir(np.random.mtrand._rand)
                 if not n.startswith('_')
                 and getattr(getattr(np.random, n, None), '__self__', None) is np.random.mtrand._rand]


class _Failure(object):
    def __init__(self, exc):
        self.exc = exc


class _Trial(object):
    def __init__(self, idx, fn, rng_state, globals_):
        self.idx, self.fn = idx, fn
        self.stream = np.random.RandomState()
        self.stream.set_state(rng_state)
        self.stream_fns = [(n, getattr(self.stream, n)) for n in _STREAM_NAMES]
        self.globals = list(globals_)
        self.go = threading.Event()
        self.request = None
        self.reply = None
        self.result = None
        self.error = None
        self.done = False
        self.thread = None
        self.pool = None


def in_flight():
    """True inside a trial thread of a running pool."""
    return getattr(_tls, 'trial', None) is not None


def merge(kind, payload, run_many):
    """Merge point.  ``run_many(list of payloads) -> list of results`` performs the device
    pass for any number of requests of this ``kind``.  Outside a pool it is called with the
    one payload; inside, the calling trial parks until the scheduler has run the merged pass."""
    trial = getattr(_tls, 'trial', None)
    if trial is None:
        return run_many([payload])[0]
    trial.request = (kind, payload, run_many)
    trial.pool._back.set()
    trial.go.wait()
    trial.go.clear()
    reply, trial.reply = trial.reply, None
    if isinstance(reply, _Failure):
        raise reply.exc
    return reply


class TrialPool(object):
    """Runs callables ("trials") in lockstep; see the module docstring.

    ``trial_globals``: (object, attribute) pairs of process-global state owned by a trial
    (saved / restored at every hand-over, each trial starting from the value at ``run``).
    ``stats`` counts the merged passes per kind and the requests they served."""

    def __init__(self, trial_globals=()):
        self.trial_globals = list(trial_globals)
        self._back = threading.Event()
        self.stats = {}
        self.seconds = {}  # wall time inside the merged passes per kind, and 'trials' (host code)

    def _enter_thread(self, trial, device, stream):
        """Thread body: CUDA's current device and torch's current stream are per thread."""
        trial.go.wait()
        trial.go.clear()
        _tls.trial = trial
        try:
            if device is not None:
                import torch
                torch.cuda.set_device(device)
                with torch.cuda.stream(stream):
                    trial.result = trial.fn()
            else:
                trial.result = trial.fn()
        except BaseException as exc:  # handed to the caller of run()
            trial.error = exc
        finally:
            _tls.trial = None
            trial.done = True
            self._back.set()

    def _resume(self, trial):
        for name, fn in trial.stream_fns:
            setattr(np.random, name, fn)
        for k, (obj, attr) in enumerate(self.trial_globals):
            setattr(obj, attr, trial.globals[k])
        self._back.clear()
        trial.go.set()
        self._back.wait()
        for k, (obj, attr) in enumerate(self.trial_globals):
            trial.globals[k] = getattr(obj, attr)

    def run(self, fns, rng_states):
        """fns: one callable per trial; rng_states: the numpy global-stream state each trial
        starts from (``np.random.get_state()`` tuples).  Returns the results in order; the
        first trial error is re-raised after all threads have ended."""
        if in_flight():
            raise RuntimeError('TrialPool.run inside a trial')
        device = stream = None
        try:
            import torch
            if torch.cuda.is_available():
                device, stream = torch.cuda.current_device(), torch.cuda.current_stream()
        except ImportError:
            pass
        own_stream = [(n, getattr(np.random, n)) for n in _STREAM_NAMES]
        own_globals = [getattr(o, a) for o, a in self.trial_globals]
        trials = [_Trial(i, fn, st, own_globals) for i, (fn, st) in enumerate(zip(fns, rng_states))]
        for t in trials:
            t.pool = self
            t.thread = threading.Thread(target=self._enter_thread, args=(t, device, stream),
                                        name='ocbo-trial-%d' % t.idx, daemon=True)
            t.thread.start()
        runnable = deque(trials)
        parked = {}
        try:
            while runnable or parked:
                if runnable:
                    t = runnable.popleft()
                    t0 = time.perf_counter()
                    self._resume(t)
                    self.seconds['trials'] = self.seconds.get('trials', 0.0) + time.perf_counter() - t0
                    if not t.done:
                        parked.setdefault(t.request[0], []).append(t)
                    continue
                kind = max(parked, key=lambda k: len(parked[k]))
                group = parked.pop(kind)
                run_many = group[0].request[2]
                n_pass, n_req = self.stats.get(kind, (0, 0))
                self.stats[kind] = (n_pass + 1, n_req + len(group))
                t0 = time.perf_counter()
                try:
                    replies = run_many([t.request[1] for t in group])
                    if len(replies) != len(group):
                        raise RuntimeError('merged pass %r returned %d results for %d requests'
                                           % (kind, len(replies), len(group)))
                except Exception as exc:
                    replies = [_Failure(exc)] * len(group)
                self.seconds[kind] = self.seconds.get(kind, 0.0) + time.perf_counter() - t0
                for t, r in zip(group, replies):
                    t.request, t.reply = None, r
                runnable.extend(group)
        finally:
            for name, fn in own_stream:
                setattr(np.random, name, fn)
            for (obj, attr), v in zip(self.trial_globals, own_globals):
                setattr(obj, attr, v)
        for t in trials:
            t.thread.join()
        for t in trials:
            if t.error is not None:
                raise t.error
        return [t.result for t in trials]


def trial_rng_states(num_trials, seed):
    """The numpy stream each in-flight trial starts from: trial 0 continues the CURRENT global
    stream (what the first trial of a sequential run sees); trial t > 0 starts from
    ``RandomState(seed + t)``.  A sequential run hands trial t whatever state trial t-1 left
    behind, which cannot be known before trial t-1 has ended -- in flight the trials' streams
    are made independent instead, and ``--trials_in_flight 1`` is the sequential run."""
    states = [np.random.get_state()]
    for t in range(1, int(num_trials)):
        states.append(np.random.RandomState((int(seed) + t) % (2 ** 32)).get_state())
    return states


def run_waves(num_trials, width, seed, make_trial, trial_globals=(), trial_ids=None):
    """``num_trials`` trials, ``width`` of them in flight at a time (they all take the same
    number of steps, so waves lose nothing).  ``make_trial(t)`` returns trial t's callable.
    ``trial_ids`` restricts the run to a rank's shard (ocbo_b200.dist.shard_trials): a trial's
    stream is keyed by its GLOBAL id, so results do not depend on the world size.
    Returns (results in the order of the ids, stats): stats[kind] = (merged passes, requests
    served), stats['seconds'] = wall time per kind and in the trials' own host code."""
    states = trial_rng_states(num_trials, seed)
    ids_all = list(range(int(num_trials))) if trial_ids is None else list(trial_ids)
    results, stats = [], {}
    seconds = stats.setdefault('seconds', {})
    for w0 in range(0, len(ids_all), int(width)):
        ids = ids_all[w0:w0 + int(width)]
        pool = TrialPool(trial_globals)
        results += pool.run([make_trial(t) for t in ids], [states[t] for t in ids])
        for k, (a, b) in pool.stats.items():
            stats[k] = (stats.get(k, (0, 0))[0] + a, stats.get(k, (0, 0))[1] + b)
        for k, v in pool.seconds.items():
            seconds[k] = seconds.get(k, 0.0) + v
    return results, stats
