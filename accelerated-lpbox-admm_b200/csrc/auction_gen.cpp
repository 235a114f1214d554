// Native batched generator of synthetic combinatorial-auction instances.
//
// Restates the PUBLISHED scheme the reference's Python generator implements -- the "arbitrary" relationships scheme of
// Leyton-Brown, Pearson & Shoham, "Towards a universal test suite for combinatorial auction algorithms" (EC-00), §4.3 --
// with the parameterisation and the operational details of generate_instances.py:137-360 (same defaults, same
// acceptance tests for substitutable bids, same dummy-item rule, same way the item-compatibility matrix is normalised
// and indexed), so that the instances have the same shape statistics as the reference's (n = n_bids variables,
// m ~ 0.38 n_bids rows, ~5.8 entries per column at j=100,k=500).  The random stream is this library's own
// (xoshiro256**, one stream per instance), NOT numpy's, so individual instances differ from the reference's files;
// the parity fixtures under tests/golden/ come from the reference generator itself.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/lpbox_b200.h"

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (int i = 0; i < 4; ++i) s[i] = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    // numpy's choice(n, p=prob): inverse-cdf search with one uniform
    int choice(const std::vector<double> &prob) {
        double u = uni(), c = 0.0;
        int last = -1;
        for (size_t i = 0; i < prob.size(); ++i) {
            if (prob[i] > 0) last = (int)i;
            c += prob[i];
            if (u < c) return (int)i;
        }
        return last < 0 ? 0 : last;
    }
};

struct Bid { std::vector<int> items; double price; };

struct Instance { int m = 0; std::vector<int> colptr, rowidx; std::vector<double> price; };

void generate(uint64_t seed, int n_items, int n_bids, double add_item_prob, Instance &out) {
    const double min_value = 1, max_value = 100, value_deviation = 0.5, additivity = 0.2, budget_factor = 1.5,
                 resale_factor = 0.5;
    const int max_n_sub_bids = 5;
    Rng rng(seed);
    std::vector<double> values(n_items);
    for (auto &v : values) v = min_value + (max_value - min_value) * rng.uni();
    // compats = triu(rand, 1); compats += compats^T; compats = compats / compats.sum(1)   (broadcast over the LAST axis)
    std::vector<double> compats((size_t)n_items * n_items, 0.0), rs(n_items, 0.0);
    for (int i = 0; i < n_items; ++i)
        for (int j = 0; j < n_items; ++j) { double u = rng.uni(); if (j > i) { compats[(size_t)i * n_items + j] = u; compats[(size_t)j * n_items + i] = u; } }
    for (int i = 0; i < n_items; ++i) { double s = 0; for (int j = 0; j < n_items; ++j) s += compats[(size_t)i * n_items + j]; rs[i] = s; }
    for (int i = 0; i < n_items; ++i) for (int j = 0; j < n_items; ++j) compats[(size_t)i * n_items + j] /= rs[j];

    std::vector<Bid> bids;
    int n_dummy = 0;
    std::vector<double> interests(n_items), pvalues(n_items), prob(n_items);
    std::vector<char> mask(n_items);
    // compats[bundle_mask, :].mean(axis=0) with an INTEGER 0/1 mask: rows 0 and 1 weighted by the mask counts
    auto choose_next = [&](int in_bundle) {
        const double w1 = (double)in_bundle / n_items, w0 = 1.0 - w1;
        double s = 0;
        for (int j = 0; j < n_items; ++j) {
            double cm = w0 * compats[j] + w1 * compats[(size_t)n_items + j];
            prob[j] = (mask[j] ? 0.0 : 1.0) * interests[j] * cm;
            s += prob[j];
        }
        for (int j = 0; j < n_items; ++j) prob[j] /= s;
        return rng.choice(prob);
    };
    auto bundle_of = [&]() { std::vector<int> b; for (int j = 0; j < n_items; ++j) if (mask[j]) b.push_back(j); return b; };
    auto price_of = [&](const std::vector<int> &b) { double p = 0; for (int j : b) p += pvalues[j]; return p + pow((double)b.size(), 1 + additivity); };
    auto resale_of = [&](const std::vector<int> &b) { double p = 0; for (int j : b) p += values[j]; return p; };

    while ((int)bids.size() < n_bids) {
        double isum = 0;
        for (int j = 0; j < n_items; ++j) { interests[j] = rng.uni(); isum += interests[j]; pvalues[j] = values[j] + max_value * value_deviation * (2 * interests[j] - 1); }
        for (int j = 0; j < n_items; ++j) prob[j] = interests[j] / isum;
        std::fill(mask.begin(), mask.end(), 0);
        int cnt = 1;
        mask[rng.choice(prob)] = 1;
        while (rng.uni() < add_item_prob) {
            if (cnt == n_items) break;
            int it = choose_next(cnt);
            if (!mask[it]) { mask[it] = 1; cnt++; }
        }
        std::vector<int> bundle = bundle_of();
        double price = price_of(bundle);
        if (price < 0) continue;
        std::vector<Bid> bidder;
        bidder.push_back({bundle, price});
        std::vector<Bid> cand;
        for (int item : bundle) {
            std::fill(mask.begin(), mask.end(), 0);
            mask[item] = 1;
            int c = 1;
            while (c < (int)bundle.size()) { int it = choose_next(c); if (!mask[it]) { mask[it] = 1; c++; } }
            std::vector<int> sb = bundle_of();
            cand.push_back({sb, price_of(sb)});
        }
        const double budget = budget_factor * price, min_resale = resale_factor * resale_of(bundle);
        std::vector<int> ord(cand.size());
        for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int)i;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return cand[a].price > cand[b].price; });
        for (int ci : ord) {
            const Bid &c = cand[ci];
            if ((int)bidder.size() >= max_n_sub_bids + 1 || (int)(bids.size() + bidder.size()) >= n_bids) break;
            if (c.price < 0 || c.price > budget) continue;
            if (resale_of(c.items) < min_resale) continue;
            bool dup = false;
            for (const Bid &b : bidder) if (b.items == c.items) { dup = true; break; }
            if (dup) continue;
            bidder.push_back(c);
        }
        int dummy = -1;
        if (bidder.size() > 2) { dummy = n_items + n_dummy; n_dummy++; }
        for (Bid &b : bidder) { if (dummy >= 0) b.items.push_back(dummy); bids.push_back(b); }
    }
    // instance_*_C.txt / _b.txt content (generate_instances.py:338-359): column = bid, rows = items (+ dummy items)
    const int n = (int)bids.size();
    out.colptr.assign(n + 1, 0); out.rowidx.clear(); out.price.resize(n);
    int max_row = -1;
    for (int i = 0; i < n; ++i) {
        std::vector<int> it = bids[i].items;
        std::sort(it.begin(), it.end());
        for (int r : it) { out.rowidx.push_back(r); max_row = std::max(max_row, r); }
        out.colptr[i + 1] = (int)out.rowidx.size();
        out.price[i] = bids[i].price;
    }
    out.m = max_row + 1;   // readSparseMat sizes the matrix by the largest row index present (LP.cpp:2438)
}

}  // namespace

extern "C" int lpbox_gen_auctions(uint64_t seed, int count, int n_items, int n_bids, double add_item_prob, int threads,
                                  int32_t **m_out, int32_t **colptr_out, int32_t **rowidx_out, double **price_out) {
    if (count <= 0 || n_items < 2 || n_bids <= 0 || !m_out || !colptr_out || !rowidx_out || !price_out) return LPBOX_E_INVALID;
    std::vector<Instance> inst(count);
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, count));
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < count;) generate(seed * 0x9e3779b97f4a7c15ULL + (uint64_t)i, n_items, n_bids, add_item_prob, inst[i]); });
    for (auto &th : pool) th.join();
    size_t tot = 0;
    for (auto &I : inst) tot += I.rowidx.size();
    int32_t *m = (int32_t *)malloc(sizeof(int32_t) * (size_t)count);
    int32_t *cp = (int32_t *)malloc(sizeof(int32_t) * (size_t)count * ((size_t)n_bids + 1));
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(tot, 1));
    double *pr = (double *)malloc(sizeof(double) * (size_t)count * (size_t)n_bids);
    if (!m || !cp || !ri || !pr) { free(m); free(cp); free(ri); free(pr); return LPBOX_E_INVALID; }
    size_t o = 0;
    for (int i = 0; i < count; ++i) {
        m[i] = inst[i].m;
        memcpy(cp + (size_t)i * (n_bids + 1), inst[i].colptr.data(), sizeof(int32_t) * ((size_t)n_bids + 1));
        memcpy(ri + o, inst[i].rowidx.data(), sizeof(int32_t) * inst[i].rowidx.size());
        memcpy(pr + (size_t)i * n_bids, inst[i].price.data(), sizeof(double) * (size_t)n_bids);
        o += inst[i].rowidx.size();
    }
    *m_out = m; *colptr_out = cp; *rowidx_out = ri; *price_out = pr;
    return 0;
}
