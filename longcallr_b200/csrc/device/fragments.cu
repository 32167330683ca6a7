/*
 * fragments.cu — read x SNP fragment matrix (CSR by read, CSC by SNP) and LD pair table.
 *
 * Replaces src/fragment.rs:10-309 (SNPFrag::get_fragments):
 *   :28-54    read filter (shared with the pileup: slot_flags) and "pos > last candidate" skip
 *   :63-80    first candidate at or after the read start
 *   :93-194   CIGAR walk; one FragElem per candidate position on an M/=/X base
 *   :207-240  allele-pair counts (only the pairs divide_snps_into_blocks can look at are kept:
 *             both SNPs for_phasing and biallelic with the reference, candidate.rs:619-676)
 *   :242-307  num_hete_links, for_phasing, snp_cover_fragments
 *
 * A cell is the int8 p * (baseq + 1) with baseq capped at 30 (fragment.rs:127-143).
 */
#include "lcr_frag.h"

namespace {

__device__ __forceinline__ bool ld_eligible(const lcr_candidate &c) {
    /* candidate.rs:640-676: for_phasing, exactly one of the two alleles is the reference, no zero frequency */
    if (!(c.flags & LCR_CF_FOR_PHASING)) return false;
    const bool a0 = c.alleles[0] == c.reference, a1 = c.alleles[1] == c.reference;
    if (a0 == a1) return false;
    return c.allele_freqs[0] != 0.0f && c.allele_freqs[1] != 0.0f;
}

/* walk one read over the region's candidates; F is called for every kept element */
template <class F>
__device__ int walk_fragment(const FragArgs &a, uint32_t read, const lcr_candidate *c, uint32_t nc, F &&emit) {
    const int64_t pos = a.pos[read];
    const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
    const uint64_t s0 = a.seq_off[read];
    const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
    const uint8_t *seq = a.seq + s0, *qual = a.qual + s0;
    uint32_t idx = 0;
    if (!(pos <= c[0].pos)) {
        uint32_t lo = 0, hi = nc;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (c[m].pos < pos) lo = m + 1; else hi = m; }
        idx = lo;
    }
    int64_t pr = pos;
    int64_t pq = lcr_leading_softclips(a.cigar, c0, c1); /* fragment.rs:59 */
    for (uint64_t ci = c0; ci < c1; ++ci) {
        const uint32_t op = a.cigar[ci], opc = op & 0xf;
        const int64_t len = op >> 4;
        if (opc == 4 || opc == 5) continue;
        if (opc == 1) { pq += len; continue; }
        const int64_t op_end = pr + len;
        if (opc == 0 || opc == 7 || opc == 8) {
            while (idx < nc && c[idx].pos < op_end) {
                const int64_t qp = pq + (c[idx].pos - pr);
                if (qp >= seq_len) return LCR_ERR_BAD_CIGAR;
                const lcr_candidate &s = c[idx];
                const uint8_t base = seq[qp];
                const uint32_t rq = qual[qp];
                const uint32_t q = rq < 30u ? rq : 30u;
                int p;
                if (base == s.reference) p = 1;
                else if (base == s.alleles[0] || base == s.alleles[1]) p = -1;
                else p = 0;
                if (!(s.flags & LCR_CF_DENSE) && p != 0) {
                    /* quality 0 (prob = 1.0, fragment.rs:133): log10(0) reaches the first sigma sweep only from a phase site, where
                       the reference panics (phase.rs:307); elsewhere the element is kept and phase.cu restates the IEEE outcomes */
                    if (q == 0 && (s.flags & LCR_CF_FOR_PHASING)) return LCR_ERR_BASEQ_ZERO;
                    emit(idx, base, (int8_t)(p * (int)(q + 1)), s);
                }
                ++idx;
            }
            pq += len;
        } else if (opc == 2 || opc == 3) {
            while (idx < nc && c[idx].pos < op_end) ++idx;
        } else return LCR_ERR_BAD_CIGAR;
        pr = op_end;
    }
    return 0;
}

__global__ void k_frag_count(FragArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_slots) return;
    uint32_t isfrag = 0, nelem = 0;
    if (a.slot_flags[slot] & 1) {
        const uint32_t reg = a.slot_region[slot];
        const LcrRegionState rs = a.rstate[reg];
        if (rs.status == 0 && rs.n_cand) {
            const lcr_candidate *c = a.cand + rs.cand_begin;
            const uint32_t read = a.regions[reg].read_begin + (slot - a.slot_off[reg]);
            if (!((int64_t)a.pos[read] > c[rs.n_cand - 1].pos)) {
                isfrag = 1;
                uint32_t elig = 0;
                const bool ld_region = rs.n_cand > a.P.max_enum_snps;
                int rc = walk_fragment(a, read, c, rs.n_cand, [&](uint32_t idx, uint8_t, int8_t, const lcr_candidate &s) {
                    nelem++;
                    atomicAdd(&a.cover_count[rs.cand_begin + idx], 1u);
                    if (ld_region && ld_eligible(s)) elig++;
                });
                if (rc) atomicMin(&a.rstate[reg].status, rc);
                if (elig > 1) atomicAdd(&a.rstate[reg].n_ld_pairs_cap, elig * (elig - 1) / 2);
            }
        }
    }
    a.frag_flag[slot] = isfrag;
    a.elem_count[slot] = nelem;
}

__global__ void k_region_frag_ranges(FragArgs a, uint32_t n_regions) {
    const uint32_t reg = blockIdx.x * blockDim.x + threadIdx.x;
    if (reg == 0) { /* the totals stay on the device; the host reads them with the counter block at the end of the run */
        a.ctr->n_frag = a.frag_scan[a.n_slots];
        a.ctr->n_elem = a.elem_scan[a.n_slots];
        if (a.elem_scan[a.n_slots] > a.elem_cap) atomicOr(&a.ctr->overflow, LCR_OVF_ELEMS);
        if (a.frag_scan[a.n_slots] == 0) a.frag_elem_off[0] = 0;
    }
    if (reg >= n_regions) return;
    const uint32_t b = a.frag_scan[a.slot_off[reg]], e = a.frag_scan[a.slot_off[reg + 1]];
    a.rstate[reg].frag_begin = b;
    a.rstate[reg].n_frag = e - b;
}

__global__ void k_frag_fill(FragArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t e0 = 0, k = 0, cb = 0, floc = 0, nnz = 0;
    const uint32_t n_frag_total = a.frag_scan[a.n_slots], n_elem_total = a.elem_scan[a.n_slots];
    if (n_elem_total > a.elem_cap) return; /* the run is repeated with larger element arrays */
    if (slot < a.n_slots && a.frag_flag[slot]) {
        const uint32_t reg = a.slot_region[slot];
        const LcrRegionState rs = a.rstate[reg];
        const uint32_t read = a.regions[reg].read_begin + (slot - a.slot_off[reg]);
        const uint32_t f = a.frag_scan[slot];
        e0 = a.elem_scan[slot];
        a.frag_slot[f] = slot;
        a.frag_elem_off[f] = e0;
        if (f + 1 == n_frag_total) a.frag_elem_off[f + 1] = n_elem_total;
        if (rs.status != 0 || !rs.n_cand) a.frag_links[f] = 0; /* a failed region reports no fragments */
        else {
            a.is_fragment[read] = 1;
            const lcr_candidate *c = a.cand + rs.cand_begin;
            uint32_t links = 0;
            cb = rs.cand_begin;
            floc = f - rs.frag_begin;
            walk_fragment(a, read, c, rs.n_cand, [&](uint32_t idx, uint8_t base, int8_t cell, const lcr_candidate &s) {
                a.elem_snp[e0 + k] = idx;
                a.elem_cell[e0 + k] = cell;
                a.elem_base[e0 + k] = base;
                ++k;
                if (s.flags & LCR_CF_FOR_PHASING) links++;
            });
            a.frag_links[f] = links;
            if (links >= a.P.min_linkers) nnz = links;
        }
    }
    nnz = __reduce_add_sync(0xffffffffu, nnz); /* one atomic per warp on the shared counter, not one per fragment */
    if (lane == 0 && nnz) atomicAdd((unsigned long long *)&a.stats->nnz_phase, (unsigned long long)nnz);
    /* cover lists (CSC): the reads of a warp are neighbours and cover the same few candidates, so the warp merges its
       element lists (each ascending by candidate) and takes one cursor atomic per candidate instead of one per element */
    uint32_t next = 0;
    for (;;) {
        const uint32_t g = next < k ? cb + a.elem_snp[e0 + next] : 0xffffffffu;
        const uint32_t gmin = __reduce_min_sync(0xffffffffu, g);
        if (gmin == 0xffffffffu) break;
        const bool mine = g == gmin;
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        const int leader = __ffs(m) - 1;
        uint32_t first = 0;
        if ((int)lane == leader) first = atomicAdd(&a.cover_cursor[gmin], (uint32_t)__popc(m));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (mine) {
            const uint32_t w = a.cover_off[gmin] + first + (uint32_t)__popc(m & ((1u << lane) - 1u));
            a.cover_frag[w] = floc;
            a.cover_cell[w] = a.elem_cell[e0 + next];
            ++next;
        }
    }
}

/* ---- the same two passes with one warp per read and one lane per CIGAR op: long CIGARs (ONT reads carry ~100 ops) ---- */

/* fragment.rs:122-152 for one candidate on an aligned base: 1 = element kept, 0 = skipped, negative = the walk's error */
__device__ __forceinline__ int eval_cand(const lcr_candidate &s, const uint8_t *seq, const uint8_t *qual, int64_t seq_len, int64_t qp, uint8_t &base, int8_t &cell) {
    if (qp >= seq_len) return LCR_ERR_BAD_CIGAR;
    base = seq[qp];
    const uint32_t rq = qual[qp];
    const uint32_t q = rq < 30u ? rq : 30u;
    int p;
    if (base == s.reference) p = 1;
    else if (base == s.alleles[0] || base == s.alleles[1]) p = -1;
    else p = 0;
    if ((s.flags & LCR_CF_DENSE) || p == 0) return 0;
    if (q == 0 && (s.flags & LCR_CF_FOR_PHASING)) return LCR_ERR_BASEQ_ZERO;
    cell = (int8_t)(p * (int)(q + 1));
    return 1;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_frag_walk_w(FragArgs a) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= a.n_slots) return;
    const uint32_t reg = a.slot_region[slot];
    const LcrRegionState rs = a.rstate[reg];
    const uint32_t read = a.regions[reg].read_begin + (slot - a.slot_off[reg]);
    const lcr_candidate *c = a.cand + rs.cand_begin;
    uint32_t isfrag = 0, f = 0, e0 = 0;
    bool go = false;
    if (!FILL) {
        if ((a.slot_flags[slot] & 1) && rs.status == 0 && rs.n_cand && !((int64_t)a.pos[read] > c[rs.n_cand - 1].pos)) { isfrag = 1; go = true; }
    } else {
        if (!a.frag_flag[slot]) return;
        if (a.elem_scan[a.n_slots] > a.elem_cap) return; /* the run is repeated with larger element arrays */
        f = a.frag_scan[slot];
        e0 = a.elem_scan[slot];
        if (lane == 0) {
            a.frag_slot[f] = slot;
            a.frag_elem_off[f] = e0;
            if (f + 1 == a.frag_scan[a.n_slots]) a.frag_elem_off[f + 1] = a.elem_scan[a.n_slots];
        }
        if (rs.status != 0 || !rs.n_cand) { if (lane == 0) a.frag_links[f] = 0; return; } /* a failed region reports no fragments */
        if (lane == 0) a.is_fragment[read] = 1;
        go = true;
    }
    uint32_t kbase = 0, links = 0, elig = 0;
    int rc = 0;
    if (go) {
        const uint32_t nc = rs.n_cand, floc = f - rs.frag_begin;
        const bool ld_region = nc > a.P.max_enum_snps;
        const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
        const uint64_t s0 = a.seq_off[read];
        const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
        const uint8_t *seq = a.seq + s0, *qual = a.qual + s0;
        long long pr = a.pos[read];
        long long pq = lcr_leading_softclips(a.cigar, c0, c1);
        for (uint64_t cbase = c0; cbase < c1; cbase += 32) {
            const uint64_t ci = cbase + lane;
            const uint32_t op = ci < c1 ? a.cigar[ci] : 4u; /* padding: a zero-length soft clip */
            const uint32_t opc = op & 0xf;
            const long long len = op >> 4;
            const bool is_m = opc == 0 || opc == 7 || opc == 8, is_dn = opc == 2 || opc == 3, is_i = opc == 1, is_sh = opc == 4 || opc == 5;
            const unsigned badmask = __ballot_sync(0xffffffffu, !(is_m || is_dn || is_i || is_sh));
            const uint32_t nvalid = badmask ? (uint32_t)__ffs(badmask) - 1u : 32u; /* the walk stops at the first unknown op */
            const bool on = lane < nvalid;
            const long long rl = (on && (is_m || is_dn)) ? len : 0, ql = (on && (is_m || is_i)) ? len : 0;
            long long rsum = rl, qsum = ql;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long r2 = __shfl_up_sync(0xffffffffu, rsum, o), q2 = __shfl_up_sync(0xffffffffu, qsum, o);
                if ((int)lane >= o) { rsum += r2; qsum += q2; }
            }
            const long long my_pr = pr + rsum - rl, my_pq = pq + qsum - ql, op_end = my_pr + rl;
            uint32_t lo_i = 0, hi_i = 0;
            if (on && is_m && len > 0) { /* candidates on [my_pr, op_end) */
                uint32_t lo = 0, hi = nc;
                while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (c[m].pos < my_pr) lo = m + 1; else hi = m; }
                lo_i = lo;
                hi = nc;
                while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (c[m].pos < op_end) lo = m + 1; else hi = m; }
                hi_i = lo;
            }
            uint32_t cnt = 0;
            int err = 0;
            for (uint32_t idx = lo_i; idx < hi_i; ++idx) {
                uint8_t base = 0;
                int8_t cell = 0;
                const lcr_candidate &sc = c[idx];
                const int r = eval_cand(sc, seq, qual, seq_len, my_pq + (sc.pos - my_pr), base, cell);
                if (r < 0) { err = r; break; }
                if (r && !FILL) {
                    atomicAdd(&a.cover_count[rs.cand_begin + idx], 1u);
                    if (ld_region && ld_eligible(sc)) elig++;
                }
                cnt += (uint32_t)r;
            }
            const unsigned errmask = __ballot_sync(0xffffffffu, err != 0);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            if (FILL) {
                uint32_t w = e0 + kbase + incl - cnt;
                for (uint32_t idx = lo_i; idx < hi_i; ++idx) {
                    uint8_t base = 0;
                    int8_t cell = 0;
                    const lcr_candidate &sc = c[idx];
                    const int r = eval_cand(sc, seq, qual, seq_len, my_pq + (sc.pos - my_pr), base, cell);
                    if (r < 0) break;
                    if (!r) continue;
                    a.elem_snp[w] = idx;
                    a.elem_cell[w] = cell;
                    a.elem_base[w] = base;
                    ++w;
                    if (sc.flags & LCR_CF_FOR_PHASING) links++;
                    const uint32_t g = rs.cand_begin + idx;
                    const uint32_t cw = a.cover_off[g] + atomicAdd(&a.cover_cursor[g], 1u);
                    a.cover_frag[cw] = floc;
                    a.cover_cell[cw] = cell;
                }
            }
            kbase += __shfl_sync(0xffffffffu, incl, 31);
            pr += __shfl_sync(0xffffffffu, rsum, 31);
            pq += __shfl_sync(0xffffffffu, qsum, 31);
            if (errmask) { rc = __shfl_sync(0xffffffffu, err, __ffs(errmask) - 1); break; }
            if (badmask) { rc = LCR_ERR_BAD_CIGAR; break; }
        }
    }
    if (!FILL) {
        elig = __reduce_add_sync(0xffffffffu, elig);
        if (lane == 0) {
            if (rc) atomicMin(&a.rstate[reg].status, rc);
            if (elig > 1) atomicAdd(&a.rstate[reg].n_ld_pairs_cap, elig * (elig - 1) / 2);
            a.frag_flag[slot] = isfrag;
            a.elem_count[slot] = kbase;
        }
    } else {
        links = __reduce_add_sync(0xffffffffu, links);
        if (lane == 0) {
            a.frag_links[f] = links;
            if (links >= a.P.min_linkers && links) atomicAdd((unsigned long long *)&a.stats->nnz_phase, (unsigned long long)links);
        }
    }
}

/* fragment.rs:207-240 restricted to the pairs candidate.rs:628-692 evaluates: cis / trans counts per SNP pair */
__global__ void k_pair_count(FragArgs a, LcrPairEntry *table) {
    if (a.ctr->pair_total == 0 || (a.ctr->overflow & (LCR_OVF_ELEMS | LCR_OVF_PAIRS))) return;
    const uint32_t n_frag_total = a.frag_scan[a.n_slots];
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n_frag_total; f += gridDim.x * blockDim.x) {
    const uint32_t slot = a.frag_slot[f];
    const uint32_t reg = a.slot_region[slot];
    const LcrRegionState rs = a.rstate[reg];
    if (rs.status != 0 || rs.n_cand <= a.P.max_enum_snps || !rs.pair_cap) continue;
    const lcr_candidate *c = a.cand + rs.cand_begin;
    const uint32_t e0 = a.frag_elem_off[f], e1 = a.frag_elem_off[f + 1];
    LcrPairEntry *tab = table + rs.pair_begin;
    const uint32_t mask = rs.pair_cap - 1;
    for (uint32_t x = e0; x < e1; ++x) {
        const uint32_t i = a.elem_snp[x];
        if (!ld_eligible(c[i])) continue;
        const int pi = a.elem_cell[x] > 0 ? 1 : -1;
        for (uint32_t y = x + 1; y < e1; ++y) {
            const uint32_t j = a.elem_snp[y];
            if (!ld_eligible(c[j])) continue;
            const int pj = a.elem_cell[y] > 0 ? 1 : -1;
            const unsigned long long key = ((unsigned long long)i << 32) | j;
            uint32_t h = (uint32_t)lcr_mix64(key) & mask;
            for (;;) {
                unsigned long long prev = atomicCAS((unsigned long long *)&tab[h].key, ~0ull, key);
                if (prev == ~0ull || prev == key) break;
                h = (h + 1) & mask;
            }
            atomicAdd(pi * pj > 0 ? &tab[h].cis : &tab[h].trans, 1u);
        }
    }
  }
}

/* candidate.rs:679-713 + snp.rs:158-188: perfect-LD pairs (min(cis, trans) == 0, |weight| >= threshold) become edges */
template <bool FILL>
__global__ void k_ld_edges(FragArgs a, const LcrPairEntry *table, const uint32_t *entry_region, uint32_t *deg, const uint32_t *adj_off, uint32_t *adj_cursor, uint32_t *adj) {
    const uint32_t ld_weight_threshold = a.P.ld_weight_threshold;
    const LcrRegionState *rstate = a.rstate;
    if (a.ctr->overflow & (LCR_OVF_ELEMS | LCR_OVF_PAIRS | LCR_OVF_ADJ)) return;
    const uint32_t table_size = a.ctr->pair_total;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < table_size; t += gridDim.x * blockDim.x) {
    const LcrPairEntry e = table[t];
    if (e.key == ~0ull) continue;
    const uint32_t c1 = e.cis < e.trans ? e.cis : e.trans, c2 = e.cis < e.trans ? e.trans : e.cis;
    if (c1 != 0 || c2 < ld_weight_threshold || c2 == 0) continue;
    const uint32_t reg = entry_region[t >> 4]; /* region of every 16-entry group (capacities are multiples of 16) */
    const uint32_t cb = rstate[reg].cand_begin;
    const uint32_t i = (uint32_t)(e.key >> 32), j = (uint32_t)e.key;
    const uint32_t sign = e.cis > e.trans ? 0u : 0x80000000u; /* weight > 0: same haplotype */
    if (!FILL) {
        atomicAdd(&deg[cb + i], 1u);
        atomicAdd(&deg[cb + j], 1u);
    } else {
        adj[adj_off[cb + i] + atomicAdd(&adj_cursor[cb + i], 1u)] = j | sign;
        adj[adj_off[cb + j] + atomicAdd(&adj_cursor[cb + j], 1u)] = i | sign;
    }
  }
}

/* GraphMap adjacency order: edges are inserted in lexicographic (i, j) order, so every node's
   neighbour list ends up ascending by neighbour index */
__global__ void k_adj_sort(FragArgs a, const uint32_t *adj_off, uint32_t *adj) {
    if (a.ctr->overflow & (LCR_OVF_ELEMS | LCR_OVF_PAIRS | LCR_OVF_ADJ)) return;
    const uint32_t n_cand = a.ctr->n_cand;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x) {
    const uint32_t b = adj_off[i], e = adj_off[i + 1];
    for (uint32_t x = b + 1; x < e; ++x) {
        const uint32_t v = adj[x];
        uint32_t y = x;
        while (y > b && (adj[y - 1] & 0x7fffffffu) > (v & 0x7fffffffu)) { adj[y] = adj[y - 1]; --y; }
        adj[y] = v;
    }
  }
}

/* adjacency size check between the two k_ld_edges passes */
__global__ void k_adj_total(FragArgs a, const uint32_t *adj_off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t t = adj_off[a.ctr->n_cand];
        a.ctr->adj_total = t;
        if (t > a.adj_cap) atomicOr(&a.ctr->overflow, LCR_OVF_ADJ);
    }
}

/* LD pair tables: one open-addressing segment per region that takes the LD path (more than max_enum_snps candidates) */
__global__ void k_pair_cap(FragArgs a, uint32_t n_regions, uint32_t *region_cap) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_regions) return;
    uint32_t cap = 0;
    if (r < n_regions) {
        const LcrRegionState s = a.rstate[r];
        if (s.status == 0 && s.n_cand > a.P.max_enum_snps && s.n_ld_pairs_cap) {
            const unsigned long long all = (unsigned long long)s.n_cand * (s.n_cand - 1) / 2;
            const unsigned long long bound = s.n_ld_pairs_cap < all ? s.n_ld_pairs_cap : all;
            unsigned long long c = 16;
            while (c < 2 * bound) c <<= 1;
            cap = c > 0x40000000ull ? 0x40000000u : (uint32_t)c;
        }
    }
    region_cap[r] = cap;
}
__global__ void k_pair_assign(FragArgs a, uint32_t n_regions, const uint32_t *region_cap, const uint32_t *region_cap_off) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = region_cap_off[n_regions];
    const bool ovf = total > a.pair_cap_total;
    if (r == 0) {
        a.ctr->pair_total = ovf ? 0u : total;
        a.ctr->pair_need = total;
        if (ovf) atomicOr(&a.ctr->overflow, LCR_OVF_PAIRS);
    }
    if (r >= n_regions) return;
    a.rstate[r].pair_begin = ovf ? 0u : region_cap_off[r];
    a.rstate[r].pair_cap = ovf ? 0u : region_cap[r];
}
__global__ void k_init_pairs(FragArgs a, LcrPairEntry *t) {
    const uint32_t n = a.ctr->pair_total;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { t[i].key = ~0ull; t[i].cis = 0; t[i].trans = 0; }
}

__global__ void k_fill_entry_region(uint32_t n_regions, const LcrRegionState *rstate, uint32_t *entry_region) {
    for (uint32_t reg = blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const LcrRegionState rs = rstate[reg];
        for (uint32_t g = threadIdx.x; g < rs.pair_cap / 16; g += blockDim.x) entry_region[rs.pair_begin / 16 + g] = reg;
    }
}

} // namespace

/* long_cigars: more than ~24 ops per read on average (ONT): one warp per read, one lane per op; else one thread per read */
void lcr_launch_frag_count(const FragArgs &a, bool long_cigars, cudaStream_t st) {
    if (!a.n_slots) return;
    if (long_cigars) k_frag_walk_w<false><<<(uint32_t)(((uint64_t)a.n_slots * 32 + 127) / 128), 128, 0, st>>>(a);
    else k_frag_count<<<(a.n_slots + 127) / 128, 128, 0, st>>>(a);
}
void lcr_launch_region_frag_ranges(const FragArgs &a, uint32_t n_regions, cudaStream_t st) {
    k_region_frag_ranges<<<(n_regions + 128) / 128, 128, 0, st>>>(a, n_regions);
}
void lcr_launch_frag_fill(const FragArgs &a, bool long_cigars, cudaStream_t st) {
    if (!a.n_slots) return;
    if (long_cigars) k_frag_walk_w<true><<<(uint32_t)(((uint64_t)a.n_slots * 32 + 127) / 128), 128, 0, st>>>(a);
    else k_frag_fill<<<(a.n_slots + 127) / 128, 128, 0, st>>>(a);
}
void lcr_launch_pair_plan(const FragArgs &a, uint32_t n_regions, uint32_t *region_cap, const uint32_t *region_cap_off, bool assign, cudaStream_t st) {
    const uint32_t g = (n_regions + 128) / 128;
    if (assign) k_pair_assign<<<g, 128, 0, st>>>(a, n_regions, region_cap, region_cap_off);
    else k_pair_cap<<<g, 128, 0, st>>>(a, n_regions, region_cap);
}
void lcr_launch_pair_build(const FragArgs &a, uint32_t n_regions, LcrPairEntry *table, uint32_t *entry_region, int sm_count, cudaStream_t st) {
    const int sms = sm_count > 0 ? sm_count : 148;
    k_init_pairs<<<sms * 4, 256, 0, st>>>(a, table);
    k_fill_entry_region<<<sms * 4, 128, 0, st>>>(n_regions, a.rstate, entry_region);
    k_pair_count<<<sms * 8, 128, 0, st>>>(a, table);
}
void lcr_launch_ld_edges(bool fill, const FragArgs &a, const LcrPairEntry *table, const uint32_t *entry_region, uint32_t *deg, const uint32_t *adj_off, uint32_t *adj_cursor,
                         uint32_t *adj, int sm_count, cudaStream_t st) {
    const int sms = sm_count > 0 ? sm_count : 148;
    if (fill) k_ld_edges<true><<<sms * 4, 256, 0, st>>>(a, table, entry_region, deg, adj_off, adj_cursor, adj);
    else k_ld_edges<false><<<sms * 4, 256, 0, st>>>(a, table, entry_region, deg, adj_off, adj_cursor, adj);
}
void lcr_launch_adj_finish(const FragArgs &a, const uint32_t *adj_off, uint32_t *adj, bool sort, int sm_count, cudaStream_t st) {
    const int sms = sm_count > 0 ? sm_count : 148;
    if (sort) k_adj_sort<<<sms * 2, 128, 0, st>>>(a, adj_off, adj);
    else k_adj_total<<<1, 32, 0, st>>>(a, adj_off);
}
