// lzb_sched.h -- host-side placement of streams on K1's persistent warps (pure C++, no CUDA).
//
// Streams are indivisible and a warp's decode speed depends on how many warps share its SM: K1 is issue-bound at
// full residency, latency-bound when few warps are resident.  Measured on B200 (tools/kbench.py, C2 streams, one
// round of n warps per SM), ms per 64 KiB stream: n=4: 17.1, 8: 18.5, 14: 22.5, 20: 24.9, 24: 27.7, 28: 30.9.
// With "longest first" alone, a batch whose sizes are spread (north-star mix: 64 KiB .. 1 MiB) ends with its longest
// streams still running: they spent the whole launch at the crowded-SM rate (measured 383 ms where the work alone
// needs ~305 ms).  The planner below therefore gives the streams that would finish late an SM with fewer resident
// warps: the first item of every warp is pre-assigned (CTA by CTA, contiguous ranges of the sorted queue), and the
// warps a CTA must not use are parked for the launch (K1 sees the LZB_ORDER_PARK sentinel and exits).  Whether to do
// so, and how aggressively, is decided by replaying the batch in a small discrete-event model of the kernel
// (tools/sched_sim.py is the Python twin used to develop it); batches of similar streams keep today's plain queue.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <functional>
#include <queue>
#include <utility>
#include <vector>

namespace lzb_sched {

static const uint32_t ORDER_PARK = 0xFFFFFFFFu;

// time per unit of work of one warp when n warps are resident on its SM, relative to n = 28
inline double rel_time(uint32_t n) {
    static const double N[] = {1, 4, 8, 14, 20, 24, 28, 64};
    static const double T[] = {17.0 / 30.9, 17.1 / 30.9, 18.5 / 30.9, 22.5 / 30.9, 24.9 / 30.9, 27.7 / 30.9, 1.0, 2.29};
    const double x = (double)n;
    if (x <= N[0]) return T[0];
    for (int i = 1; i < 8; i++)
        if (x <= N[i]) return T[i - 1] + (T[i] - T[i - 1]) * (x - N[i - 1]) / (N[i] - N[i - 1]);
    return T[7];
}

// Replays a launch.  work[k] = work of queue entry k (any unit).  counts == nullptr: plain queue, `sms` CTAs of
// `warps` warps, initial entries taken round-robin.  counts != nullptr: CTA c owns the next counts[c] queue entries
// as its warps' first items and keeps counts[c] warps for the whole launch; the rest of the queue is dynamic.
// Returns the makespan in units of (work unit x time per unit at full residency).
inline double simulate(const std::vector<double>& work, const std::vector<uint32_t>* counts, uint32_t sms, uint32_t warps) {
    struct Sm {
        std::priority_queue<double, std::vector<double>, std::greater<double>> fin;  // virtual finish marks
        double v = 0, t = 0;  // work done per resident stream so far; wall time of the last update
        uint32_t cap = 0;
    };
    const size_t n = work.size();
    const uint32_t ctas = counts ? (uint32_t)counts->size() : sms;
    std::vector<Sm> sm(ctas);
    size_t nxt = 0;
    if (counts) {
        for (uint32_t c = 0; c < ctas; c++) {
            sm[c].cap = (*counts)[c];
            for (uint32_t k = 0; k < (*counts)[c] && nxt < n; k++) sm[c].fin.push(work[nxt++]);
        }
    } else {
        for (uint32_t c = 0; c < ctas; c++) sm[c].cap = warps;
        for (uint32_t r = 0; r < warps && nxt < n; r++)
            for (uint32_t c = 0; c < ctas && nxt < n; c++) sm[c].fin.push(work[nxt++]);
    }
    typedef std::pair<double, uint32_t> Ev;  // (time of the CTA's next completion, cta)
    std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> heap;
    auto arm = [&](uint32_t c) {
        Sm& s = sm[c];
        if (!s.fin.empty()) heap.push(Ev(s.t + (s.fin.top() - s.v) * rel_time((uint32_t)s.fin.size()), c));
    };
    for (uint32_t c = 0; c < ctas; c++) arm(c);
    double end = 0;
    while (!heap.empty()) {
        const Ev e = heap.top();
        heap.pop();
        Sm& s = sm[e.second];
        s.v = s.fin.top();  // every resident stream advanced by the same amount
        s.t = e.first;
        end = std::max(end, s.t);
        while (!s.fin.empty() && s.fin.top() <= s.v + 1e-12) s.fin.pop();
        while (nxt < n && s.fin.size() < s.cap) s.fin.push(s.v + work[nxt++]);
        arm(e.second);
    }
    return end;
}

struct Plan {
    bool throttled = false;        // false: plain queue (order = the input order, n_static = 0)
    std::vector<uint32_t> order;   // [n_static pre-assigned first items, ORDER_PARK for parked warps][dynamic queue]
    uint32_t n_static = 0, grid = 0, parked = 0;
    double predicted = 0, plain = 0;  // model makespans (same unit as simulate)
};

// sorted: stream indices, longest first; work[i] = work of stream i (K1 time is ~proportional to compressed bytes).
inline Plan plan(const std::vector<uint32_t>& sorted, const std::vector<double>& work_of, uint32_t sms, uint32_t warps) {
    Plan best;
    const size_t n = sorted.size();
    best.order = sorted;
    best.grid = (uint32_t)std::max<size_t>(1, std::min<size_t>(sms, (n + warps - 1) / warps));
    // only batches of more than one round can lose time to a tail, and parking needs residency to give away
    if (n <= (size_t)sms * warps || warps < 8) return best;
    std::vector<double> w(n);
    double total = 0;
    for (size_t k = 0; k < n; k++) total += (w[k] = std::max(1.0, work_of[sorted[k]]));
    best.plain = best.predicted = simulate(w, nullptr, sms, warps);
    const double ideal = total / ((double)sms * warps);
    if (best.plain <= ideal * 1.04) return best;  // nothing to gain
    static const double alphas[] = {1.03, 1.06, 1.09, 1.12, 1.16, 1.22};
    const uint32_t nmin = 4;
    for (double alpha : alphas) {
        const double D = alpha * std::max(ideal, w[0] * rel_time(1));
        // nlim[k]: most resident warps under which stream k, started at t = 0, still ends by D
        std::vector<uint32_t> counts;
        size_t nxt = 0;
        uint32_t parked = 0;
        for (uint32_t c = 0; c < sms && nxt < n; c++) {
            uint32_t cnt = 0, lim = warps;
            while (nxt < n && cnt < warps) {
                const double need = D / w[nxt];
                uint32_t nl = warps;
                while (nl > nmin && rel_time(nl) > need) nl--;
                const uint32_t l2 = std::min(lim, nl);
                if (cnt + 1 > l2) break;
                lim = l2;
                cnt++;
                nxt++;
            }
            counts.push_back(cnt);
            parked += warps - cnt;
        }
        if (parked == 0 || parked > (uint32_t)(0.15 * sms * warps)) continue;
        const double t = simulate(w, &counts, sms, warps);
        // the first taker must beat the plain queue by a margin the model can be trusted with
        if (t < (best.throttled ? best.predicted : best.plain * 0.97)) {
            best.predicted = t;
            best.throttled = true;
            best.parked = parked;
            best.grid = (uint32_t)counts.size();
            best.n_static = best.grid * warps;
            best.order.clear();
            size_t k = 0;
            for (uint32_t c = 0; c < best.grid; c++) {
                for (uint32_t j = 0; j < warps; j++) best.order.push_back(j < counts[c] ? sorted[k++] : ORDER_PARK);
            }
            for (; k < n; k++) best.order.push_back(sorted[k]);
        }
    }
    return best;
}

}  // namespace lzb_sched
