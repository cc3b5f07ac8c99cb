// removeDup (cLoops/cModel.py:198-259) for the loops of one chromosome, host C++ behind the C ABI.
//
// The reference walks the loops in key order: loop i (never the last one, :207) that is not yet grouped becomes the
// leader of a group holding every later, not yet grouped loop j whose two anchors both overlap i's (checkOverlap,
// :185-195, closed intervals); loops without any such j are unique.  Then every group keeps ONE member: among those with
// binomial p <= bpcut the one with the largest rab/ra/rb (:235-258).  Output order = the unique loops in key order, then
// the group winners in leader order (dict insertion order), which the second removeDup pass and the final table inherit.
//
// The O(n^2) pair loop becomes a sweep: a loop j can only overlap i if its left anchor starts inside
// [a0_i - widest left anchor, a1_i], so candidates come from an a0-sorted index; grouping stays the reference's greedy,
// order-dependent rule (the leader is the FIRST key, members join in key order).  When the maximum of a group is shared by
// several members the reference's winner depends on the sort pandas uses; such groups are handed back to the caller
// (tie lists), which resolves them with the reference's own expression.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {
inline bool one_end(long long xa, long long xb, long long ya, long long yb) {      // checkOneEndOverlap, cModel.py:174-182
    return (ya <= xa && xa <= yb) || (ya <= xb && xb <= yb) || (xa <= ya && ya <= xb) || (xa <= yb && yb <= xb);
}
}  // namespace

extern "C" int cloops_remove_dup(const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1, const double* bp,
                                 const double* dens, int64_t n, double bpcut, int64_t* keep, int64_t* n_keep, int64_t* tie_start,
                                 int64_t* tie_members, int64_t* n_ties) {
    using cloops::fail;
    if (n < 0 || !keep || !n_keep || !tie_start || !tie_members || !n_ties) return fail(CLOOPS_EINVAL, "bad argument");
    *n_keep = 0;
    *n_ties = 0;
    tie_start[0] = 0;
    if (n == 0) return 0;
    bool proper = true;
    long long wmax = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (a0[i] > a1[i] || b0[i] > b1[i]) proper = false;
        wmax = std::max<long long>(wmax, a1[i] - a0[i]);
    }
    std::vector<int64_t> order(n);
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    if (proper) std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return a0[x] < a0[y]; });
    std::vector<long long> a0s(n);
    for (int64_t k = 0; k < n; ++k) a0s[k] = a0[order[k]];
    std::vector<char> taken(n, 0);
    std::vector<int64_t> uniq, leaders, gstart(1, 0), gmem, cand;
    for (int64_t i = 0; i + 1 < n; ++i) {                      // the last key never leads (and is dropped unless grouped), :207
        if (taken[i]) continue;
        cand.clear();
        if (proper) {
            const int64_t lo = std::lower_bound(a0s.begin(), a0s.end(), (long long)a0[i] - wmax) - a0s.begin();
            const int64_t hi = std::upper_bound(a0s.begin(), a0s.end(), (long long)a1[i]) - a0s.begin();
            for (int64_t k = lo; k < hi; ++k) {
                const int64_t j = order[k];
                if (j > i && !taken[j] && one_end(a0[i], a1[i], a0[j], a1[j]) && one_end(b0[i], b1[i], b0[j], b1[j])) cand.push_back(j);
            }
            std::sort(cand.begin(), cand.end());
        } else {
            for (int64_t j = i + 1; j < n; ++j)
                if (!taken[j] && one_end(a0[i], a1[i], a0[j], a1[j]) && one_end(b0[i], b1[i], b0[j], b1[j])) cand.push_back(j);
        }
        if (cand.empty()) { uniq.push_back(i); continue; }
        leaders.push_back(i);
        taken[i] = 1;
        gmem.push_back(i);
        for (int64_t j : cand) { taken[j] = 1; gmem.push_back(j); }
        gstart.push_back((int64_t)gmem.size());
    }
    int64_t nk = 0, nt = 0, tm = 0;
    for (int64_t u : uniq) keep[nk++] = u;
    for (size_t g = 0; g < leaders.size(); ++g) {
        int64_t best = -1, n_best = 0;
        for (int64_t t = gstart[g]; t < gstart[g + 1]; ++t) {
            const int64_t m = gmem[t];
            if (bp[m] > bpcut) continue;                       // NaN compares false, as in the reference
            if (best < 0 || dens[m] > dens[best]) { best = m; n_best = 1; }
            else if (dens[m] == dens[best]) ++n_best;
            else if (dens[m] != dens[m] || dens[best] != dens[best]) n_best = 2;     // NaN: leave the order to the caller
        }
        if (best < 0) continue;                                // no significant member: the group vanishes (:251-252)
        if (n_best == 1) { keep[nk++] = best; continue; }
        keep[nk++] = -(nt + 1);                                // placeholder: tie group nt, resolved by the caller
        for (int64_t t = gstart[g]; t < gstart[g + 1]; ++t)
            if (!(bp[gmem[t]] > bpcut)) tie_members[tm++] = gmem[t];
        tie_start[++nt] = tm;
    }
    *n_keep = nk;
    *n_ties = nt;
    return 0;
}

// combineTwice (cLoops/pipe.py:155-174) over all clustering rounds of one chromosome at once.  rows int32[n,4] = the
// inter-ligation candidates (minX, maxX, minY, maxY) of every round, concatenated in round order; round int32[n] = the
// round each row came from (non-decreasing).  keep[i] = 0 iff the exact same box was already produced by an EARLIER round
// (boxes repeated inside one round are all kept, as the reference keeps them).  One pass over an open-addressing table.
extern "C" int cloops_combine_rounds(const int32_t* rows, const int32_t* round, int64_t n, uint8_t* keep) {
    using cloops::fail;
    if (n < 0 || (n > 0 && (!rows || !round || !keep))) return fail(CLOOPS_EINVAL, "bad argument");
    if (n == 0) return 0;
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int64_t> slot(cap, -1);                        // index of the first row holding the box
    const size_t mask = cap - 1;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* r = rows + 4 * i;
        uint64_t h = ((uint64_t)(uint32_t)r[0] << 32 | (uint32_t)r[1]) * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)r[2] << 32 | (uint32_t)r[3]) * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
        size_t at = (size_t)h & mask;
        for (;;) {
            const int64_t j = slot[at];
            if (j < 0) { slot[at] = i; keep[i] = 1; break; }
            const int32_t* q = rows + 4 * j;
            if (q[0] == r[0] && q[1] == r[1] && q[2] == r[2] && q[3] == r[3]) { keep[i] = round[j] == round[i]; break; }
            at = (at + 1) & mask;
        }
    }
    return 0;
}
