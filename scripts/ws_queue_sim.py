"""Discrete-event model of the warp-specialised kernel's request queues (development aid).

Controller groups post rounds of requests (one per lane), worker warps pull items; a group resumes only when its whole
round is served, runs `scalar` cycles of solver code, and posts the next round.  Used to compare pull policies and
group shapes before spending GPU time.  Numbers: cycles measured with -DPNJL_PROFILE_PHASES on cfg5."""
import heapq
import random
import sys


def simulate(groups, n_workers=14, policy="alternate", scalar=28000, t_fj=45000, t_ft=67000, p_ft=0.34, p_exit=0.09,
             horizon=60e6, jitter=0.1, seed=0):
    rng = random.Random(seed)
    ng = len(groups)
    unpulled = [[] for _ in range(ng)]     # per group: list of item costs not yet handed out
    pending = [0] * ng                     # items of the open round not yet finished
    posted_at = [0.0] * ng
    events = []                            # (time, kind, payload)
    free_workers = list(range(n_workers))
    last = [0] * n_workers                 # alternate policy: group index to look at next
    busy = 0.0
    t = 0.0

    def post(g, now):
        items = []
        for _ in range(groups[g]):
            u = rng.random()
            if u < p_exit:
                c = 0.0
            else:
                c = (t_ft if rng.random() < p_ft else t_fj) * (1 + jitter * (rng.random() - 0.5))
            items.append(c)
        unpulled[g] = items
        pending[g] = len(items)
        posted_at[g] = now

    def pull(w):
        order = list(range(ng))
        if policy == "alternate":
            order = order[last[w]:] + order[:last[w]]
        elif policy == "oldest":
            order.sort(key=lambda g: posted_at[g])
        for g in order:
            if unpulled[g]:
                c = unpulled[g].pop()
                if policy == "alternate":
                    last[w] = (g + 1) % ng
                return g, c
        return None

    for g in range(ng):
        post(g, 0.0)
    waits = []
    while t < horizon:
        # hand out work to free workers
        progressed = True
        while free_workers and progressed:
            progressed = False
            w = free_workers[-1]
            got = pull(w)
            if got is not None:
                free_workers.pop()
                g, c = got
                busy += c
                heapq.heappush(events, (t + c, 0, (w, g)))
                progressed = True
        if not events:
            break
        t, kind, payload = heapq.heappop(events)
        if kind == 0:
            w, g = payload
            free_workers.append(w)
            pending[g] -= 1
            if pending[g] == 0:
                waits.append(t - posted_at[g])
                heapq.heappush(events, (t + scalar, 1, g))
        else:
            post(payload, t)
    return busy / (n_workers * t), sum(waits) / len(waits)


if __name__ == "__main__":
    for name, groups, nw in (("2x28, 14 workers", [28, 28], 14), ("4x14, 12 workers", [14] * 4, 12), ("4x14, 14 workers", [14] * 4, 14),
                             ("32+24, 14 workers", [32, 24], 14), ("3x19, 13 workers", [19, 19, 18], 13),
                             ("8x7, 14 workers", [7] * 8, 14), ("56x1, 14 workers", [1] * 56, 14)):
        for pol in ("alternate", "priority", "oldest"):
            u, wt = simulate(groups, nw, pol)
            print("%-22s %-10s worker busy %.3f  mean round wait %.0f" % (name, pol, u, wt))
