#!/usr/bin/env python3
"""tools/sched_sim.py -- discrete-event model of K1's persistent-warp scheduling (development aid, not shipped).

Per-warp decode speed depends on how many warps are active on the SM.  Measured on B200 (tools/kbench.py, C2 streams,
one round of n warps per SM): ms per 64 KiB stream = {4: 17.1, 8: 18.5, 14: 22.5, 20: 24.9, 24: 27.7, 28: 30.9}.
The model replays a batch (work per stream in 64 KiB units) under an admission policy and reports the makespan.
"""
import heapq
import sys

import numpy as np

PTS_N = np.array([1, 4, 8, 14, 20, 24, 28], dtype=float)
PTS_T = np.array([17.0, 17.1, 18.5, 22.5, 24.9, 27.7, 30.9])


def t_unit(n):
    return float(np.interp(n, PTS_N, PTS_T))


def simulate(work, cost=None, sms=148, slots=28, cap=28.0):
    """work: per-stream work, already in queue order.  cost: per-stream admission cost (units of `cap` per SM); a warp
    takes the next stream only if the SM's used units + cost <= cap (else it idles until a stream on its SM ends)."""
    n = len(work)
    cost = np.ones(n) if cost is None else cost
    nxt = 0
    now = 0.0
    # per SM: list of remaining work of active streams, used units
    rem = [dict() for _ in range(sms)]   # stream id -> remaining work
    used = [0.0] * sms
    busy_integral = 0.0

    def admit(s):
        nonlocal nxt
        while nxt < n and len(rem[s]) < slots and used[s] + cost[nxt] <= cap + 1e-9:
            rem[s][nxt] = work[nxt]
            used[s] += cost[nxt]
            nxt += 1

    # initial fill: round-robin over SMs like warps racing for the counter
    progress = True
    while progress and nxt < n:
        progress = False
        for s in range(sms):
            if nxt < n and len(rem[s]) < slots and used[s] + cost[nxt] <= cap + 1e-9:
                rem[s][nxt] = work[nxt]
                used[s] += cost[nxt]
                nxt += 1
                progress = True
    # event loop: per SM, time to next completion = min(rem) * t_unit(n_active)
    last = [0.0] * sms
    heap = []
    ver = [0] * sms

    def push(s):
        if rem[s]:
            tu = t_unit(len(rem[s]))
            heapq.heappush(heap, (last[s] + min(rem[s].values()) * tu, s, ver[s]))

    for s in range(sms):
        push(s)
    end = 0.0
    while heap:
        t, s, v = heapq.heappop(heap)
        if v != ver[s]:
            continue
        tu = t_unit(len(rem[s]))
        done = (t - last[s]) / tu
        fin = []
        for k in list(rem[s]):
            rem[s][k] -= done
            if rem[s][k] <= 1e-9:
                fin.append(k)
        for k in fin:
            used[s] -= cost[k]
            del rem[s][k]
        last[s] = t
        end = max(end, t)
        admit(s)
        ver[s] += 1
        push(s)
    return end


def ns_sizes(n=8192, distinct=4096, index=6):
    out = []
    for i in range(distinct):
        seed = index * 1_000_003 + i
        out.append(int(65536 * 16 ** np.random.default_rng(seed ^ 0x5EED).random()))
    return np.array([out[i % distinct] for i in range(n)], dtype=float)


def costs_for(work, alpha, sms=148, slots=28):
    """Admission cost of each stream: slots / n_i, n_i = the most co-resident warps under which the stream still ends by
    the deadline D = alpha * (ideal makespan)."""
    ideal = work.sum() * t_unit(slots) / (sms * slots)
    D = alpha * max(ideal, work.max() * t_unit(1))
    cost = np.ones(len(work))
    for i, w in enumerate(work):
        need = D / w  # ms per unit this stream may take
        if need >= t_unit(slots):
            continue
        ni = slots
        for cand in range(slots, 0, -1):
            if t_unit(cand) <= need:
                ni = cand
                break
        else:
            ni = 1
        cost[i] = slots / max(ni, 4)
    return cost, D, ideal


if __name__ == "__main__":
    sizes = ns_sizes()
    work = np.sort(sizes / 65536.0)[::-1]
    print("streams", len(work), "total units", work.sum(), "max", work.max())
    base = simulate(work)
    print(f"current policy (longest first, no admission control): {base:.1f} ms   (measured on B200: 383 ms)")
    for alpha in (1.0, 1.03, 1.06, 1.1, 1.15, 1.2, 1.3):
        cost, D, ideal = costs_for(work, alpha)
        t = simulate(work, cost)
        print(f"alpha {alpha:.2f}: deadline {D:.0f} (ideal {ideal:.0f})  throttled streams {(cost > 1).sum():5d}  makespan {t:.1f} ms")


def nlimit_for(work, D, slots=28, nmin=4):
    """n_i = most co-resident warps under which stream i still ends by D when started at t = 0."""
    out = np.full(len(work), slots, dtype=int)
    for i, w in enumerate(work):
        need = D / w
        ni = nmin
        for cand in range(slots, nmin - 1, -1):
            if t_unit(cand) <= need:
                ni = cand
                break
        out[i] = ni
    return out


def simulate_nlimit(work, nlim, sms=148, slots=28):
    """Queue sorted by work (desc) => nlim non-decreasing.  Initial fill: SM by SM, contiguous queue ranges, each SM
    takes streams while count < min(nlim of its streams).  Afterwards a free warp takes the next stream iff
    active < min(nlim of the SM's active streams)."""
    n = len(work)
    nxt = 0
    rem = [dict() for _ in range(sms)]
    for s in range(sms):
        while nxt < n and len(rem[s]) < slots:
            lim = min([nlim[k] for k in rem[s]] + [nlim[nxt]])
            if len(rem[s]) + 1 > lim:
                break
            rem[s][nxt] = work[nxt]
            nxt += 1
    last = [0.0] * sms
    ver = [0] * sms
    heap = []

    def push(s):
        if rem[s]:
            heapq.heappush(heap, (last[s] + min(rem[s].values()) * t_unit(len(rem[s])), s, ver[s]))

    for s in range(sms):
        push(s)
    end = 0.0
    while heap:
        t, s, v = heapq.heappop(heap)
        if v != ver[s]:
            continue
        done = (t - last[s]) / t_unit(len(rem[s]))
        for k in list(rem[s]):
            rem[s][k] -= done
            if rem[s][k] <= 1e-9:
                del rem[s][k]
        last[s] = t
        end = max(end, t)
        while nxt < n and len(rem[s]) < slots:
            lim = min([nlim[k] for k in rem[s]] + [slots])
            if len(rem[s]) + 1 > lim:
                break
            rem[s][nxt] = work[nxt]
            nxt += 1
        ver[s] += 1
        push(s)
    return end


if __name__ == "__main__":
    ideal = work.sum() * t_unit(28) / (148 * 28)
    for alpha in (0.95, 1.0, 1.03, 1.06, 1.1, 1.15, 1.2):
        D = alpha * max(ideal, work.max() * t_unit(1))
        nl = nlimit_for(work, D)
        t = simulate_nlimit(work, nl)
        print(f"n-limit policy alpha {alpha:.2f}: D {D:.0f}  throttled {(nl < 28).sum():5d}  min n {nl.min()}  makespan {t:.1f} ms")


def simulate_classes(work, nlim, bounds, sms=148, slots=28):
    """Multi-launch policy: streams with nlim < slots are split into classes at the given nlim boundaries; class c runs
    as its own launch with n_c = min nlim of the class warps per CTA, one stream per warp (all start at t = 0), one CTA
    per SM.  Everything else is the main launch (slots warps per CTA) on the remaining SMs; an SM joins the main launch
    when its class CTA has exited (all its warps done)."""
    n = len(work)
    idx_thr = [i for i in range(n) if nlim[i] < slots]
    classes = []
    lo = 0
    for b in bounds + [slots]:
        members = [i for i in idx_thr if lo <= nlim[i] < b]
        lo = b
        if members:
            classes.append(members)
    sm_free_at = []
    for members in classes:
        nc = int(min(nlim[i] for i in members))
        for g in range(0, len(members), nc):
            grp = [work[i] for i in members[g:g + nc]]
            # streams of one CTA progress together; rate changes as they finish
            rem = sorted(grp)
            t = 0.0
            donew = 0.0
            k = len(rem)
            for j, w in enumerate(rem):
                t += (w - donew) * t_unit(k - j)
                donew = w
            sm_free_at.append(t)
    if len(sm_free_at) > sms:
        return float("inf"), len(sm_free_at)
    main = [i for i in range(n) if nlim[i] >= slots]
    mwork = [work[i] for i in main]
    # main launch: SMs available at time 0 (sms - X) or at sm_free_at
    avail = sorted([0.0] * (sms - len(sm_free_at)) + sm_free_at)
    nxt = 0
    rem = [dict() for _ in range(sms)]
    last = list(avail)
    ver = [0] * sms
    heap = []
    started = [False] * sms
    for s in range(sms):
        heapq.heappush(heap, (avail[s], s, -1))
    end = max(sm_free_at) if sm_free_at else 0.0
    while heap:
        t, s, v = heapq.heappop(heap)
        if v == -1:
            started[s] = True
        elif v != ver[s]:
            continue
        if rem[s]:
            done = (t - last[s]) / t_unit(len(rem[s]))
            for k in list(rem[s]):
                rem[s][k] -= done
                if rem[s][k] <= 1e-9:
                    del rem[s][k]
            end = max(end, t)
        last[s] = t
        while nxt < len(mwork) and len(rem[s]) < slots:
            rem[s][nxt] = mwork[nxt]
            nxt += 1
        ver[s] += 1
        if rem[s]:
            heapq.heappush(heap, (t + min(rem[s].values()) * t_unit(len(rem[s])), s, ver[s]))
    return end, len(sm_free_at)


if __name__ == "__main__":
    for alpha in (1.0, 1.05, 1.1, 1.15):
        D = alpha * max(ideal, work.max() * t_unit(1))
        nl = nlimit_for(work, D)
        for bounds in ([], [14], [12, 16, 20, 24], [10, 12, 14, 16, 18, 20, 22, 24, 26]):
            t, x = simulate_classes(work, nl, bounds)
            print(f"class launches alpha {alpha:.2f} bounds {bounds}: class SMs {x:3d}  makespan {t:.1f} ms")


def simulate_park(work, nlim, sms=148, slots=28):
    """Static variant: the first item of every warp is pre-assigned (SM by SM, contiguous ranges of the sorted queue, an
    SM takes streams while count < min nlim of its streams); its other warps are parked for the whole launch.  The
    rest of the queue is pulled dynamically by the non-parked warps."""
    n = len(work)
    nxt = 0
    rem = [dict() for _ in range(sms)]
    cap = [slots] * sms
    for s in range(sms):
        while nxt < n and len(rem[s]) < slots:
            lim = min([nlim[k] for k in rem[s]] + [nlim[nxt]])
            if len(rem[s]) + 1 > lim:
                break
            rem[s][nxt] = work[nxt]
            nxt += 1
        cap[s] = max(1, len(rem[s])) if nxt < n or rem[s] else slots
    last = [0.0] * sms
    ver = [0] * sms
    heap = []
    for s in range(sms):
        if rem[s]:
            heapq.heappush(heap, (min(rem[s].values()) * t_unit(len(rem[s])), s, 0))
    end = 0.0
    while heap:
        t, s, v = heapq.heappop(heap)
        if v != ver[s]:
            continue
        done = (t - last[s]) / t_unit(len(rem[s]))
        for k in list(rem[s]):
            rem[s][k] -= done
            if rem[s][k] <= 1e-9:
                del rem[s][k]
        last[s] = t
        end = max(end, t)
        while nxt < n and len(rem[s]) < cap[s]:
            rem[s][nxt] = work[nxt]
            nxt += 1
        ver[s] += 1
        if rem[s]:
            heapq.heappush(heap, (t + min(rem[s].values()) * t_unit(len(rem[s])), s, ver[s]))
    return end, sum(slots - c for c in cap)


if __name__ == "__main__":
    for alpha in (1.0, 1.05, 1.1, 1.15, 1.2):
        D = alpha * max(ideal, work.max() * t_unit(1))
        nl = nlimit_for(work, D)
        t, parked = simulate_park(work, nl)
        print(f"static park alpha {alpha:.2f}: parked warps {parked:4d}  makespan {t:.1f} ms")
