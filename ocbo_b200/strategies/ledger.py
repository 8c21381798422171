"""Per-run bookkeeping shared by the discrete strategies.

The reference keeps its histories as loose attributes of the strategy object
(``query_history, curr_best, total_best, choice_timeline, rn_choice, rn_total,
allocation_count``; /root/reference/src/strategies/multi_opt.py:169-176,178-193)
and its drivers / plotters read exactly those names.  ``TaskLedger`` owns the
containers and the update rules; the strategy re-exports the containers under
the reference's attribute names.
"""
from argparse import Namespace


class TaskLedger(object):

    def __init__(self, task_names, track_risk_neutral):
        self.names = list(task_names)
        self.track_rn = bool(track_risk_neutral)
        self.query_history = {n: [] for n in self.names}   # per task: Namespace(pt, val, init_pt, time)
        self.curr_best = {n: [] for n in self.names}       # per task: running (val, pt) incumbent
        self.rn_choice = {n: [] for n in self.names}       # per task: risk-neutral (val, pt)
        self.total_best = []                               # sum of incumbents after every step
        self.rn_total = []
        self.choice_timeline = []                          # task index chosen at every step
        self.allocation_count = [0] * len(self.names)

    def observations(self, f_idx):
        log = self.query_history[self.names[f_idx]]
        return [q.pt for q in log], [q.val for q in log]

    def incumbent(self, f_idx):
        return self.curr_best[self.names[f_idx]][-1][0]

    def num_queries(self, f_idx):
        return len(self.query_history[self.names[f_idx]])

    def record(self, f_idx, q_info, step, rn_entry=None):
        """Append one evaluation.  ``rn_entry`` is the risk-neutral (value, point) for
        this step (the observation itself for initial points).  Totals and the choice
        timeline only start once the optimisation proper is running (step > 0)."""
        name = self.names[f_idx]
        self.allocation_count[f_idx] += 1
        self.query_history[name].append(q_info)
        if self.track_rn:
            self.rn_choice[name].append(rn_entry)
        trail = self.curr_best[name]
        if trail and not trail[-1][0] < q_info.val:
            trail.append(trail[-1])
        else:
            trail.append((q_info.val, q_info.pt))
        if step > 0:
            acc = 0
            for t in self.curr_best.values():
                acc += t[-1][0]
            self.total_best.append(acc)
            self.choice_timeline.append(f_idx)
            if self.track_rn:
                acc = 0
                for t in self.rn_choice.values():
                    acc += t[-1][0]
                self.rn_total.append(acc)

    def as_histories(self):
        return Namespace(query_history=self.query_history, curr_best=self.curr_best,
                         total_best=self.total_best, choice_timeline=self.choice_timeline,
                         rn_choice=self.rn_choice, rn_total=self.rn_total)
