/*
 * lcr_pipeline.h — scratch layout of one device run (internal).
 *
 * lcr_run_device issues every kernel of the path on the context stream without a host round trip: every
 * data-dependent size (items, segments, pre-candidates, fragment elements, LD pair tables, enumeration work lists)
 * lives in device counters, the buffers are carved from a per-context arena by capacities kept with the batch handle,
 * and kernels that would exceed a capacity set an overflow bit instead of writing.  The host reads the counter block
 * once at the end of the run; on overflow it grows the capacities to what the counters ask for and runs again.
 */
#ifndef LCR_PIPELINE_H
#define LCR_PIPELINE_H

#include <stddef.h>
#include <stdint.h>

#define LCR_ENUM_SHAPES 5
#define LCR_ENUM_CLASSES 4
#define LCR_ENUM_BINS (LCR_ENUM_SHAPES * LCR_ENUM_CLASSES)

/* capacity overflow bits (LcrCounters::overflow) */
#define LCR_OVF_ITEMS 1u
#define LCR_OVF_SEGS 2u
#define LCR_OVF_PRE 4u
#define LCR_OVF_ELEMS 8u
#define LCR_OVF_PAIRS 16u
#define LCR_OVF_ADJ 32u
#define LCR_OVF_ENUM 64u

struct LcrCaps {
    uint64_t items, segs, pre, elems, pairs, adj, enum_work;
};

/* device counters of one run: zeroed when the run starts, copied to the host when it ends */
struct LcrCounters {
    uint32_t overflow;
    uint32_t n_items;        /* item slots reserved (sum of the per-tile upper bounds) */
    uint32_t n_segs;         /* segment slots reserved (sum of the per-read upper bounds) */
    uint32_t n_pre;          /* sites that passed the count filters */
    uint32_t n_cand;
    uint32_t n_frag, n_elem;
    uint32_t pair_total;     /* LD pair table entries in use */
    uint32_t adj_total;
    uint32_t enum_work;      /* enumeration work items */
    uint32_t n_list[2];      /* tiles in the shallow / deep work list */
    uint32_t ticket[2];      /* their dynamic schedulers */
    uint32_t n_tiles_done;   /* tiles the pileup kernel processed */
    uint32_t n_segs_used;    /* segments actually written */
    uint32_t n_items_used;   /* items actually registered */
    uint32_t pad0;
    unsigned long long big_cfgs; /* configurations of the 5+ site enumeration regions (launch-shape choice) */
    unsigned long long n_pos_done; /* positions of the processed tiles */
    unsigned long long qual_reads; /* base qualities k_site_ll fetched (accounting of LCR_FLAG_QUAL_ON_DEMAND) */
    uint32_t enum_cnt[LCR_ENUM_BINS], enum_off[LCR_ENUM_BINS], enum_ticket[LCR_ENUM_BINS];
    uint32_t n_big;          /* LD-path regions handed to the cooperative kernel */
    uint32_t pair_need;      /* LD pair table entries the batch asks for */
    unsigned long long prof[16]; /* LCR_TILE_PROF builds: cycles per phase of the tile kernel */
};

/* bump allocator over one device block */
struct LcrArena {
    char *base = nullptr;
    size_t cap = 0, off = 0;
    bool dry = false; /* sizing pass: only add up */
    template <class T>
    T *take(size_t n) {
        const size_t bytes = (sizeof(T) * (n ? n : 1) + 255) & ~(size_t)255;
        T *p = dry ? nullptr : reinterpret_cast<T *>(base + off);
        off += bytes;
        return p;
    }
};

#endif
