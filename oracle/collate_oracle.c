/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Plain-C restatement of the two native ops of the reference's `geotransformer.ext`:
 *
 *   oracle_grid_subsampling  <- geotransformer/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp:3-75
 *                               (+ extra/cloud/cloud.cpp:4-37 min/max corner, cloud.h:84-98 operators)
 *   oracle_radius_neighbors  <- geotransformer/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91
 *                               (+ vendored nanoflann v0x130: L2_Simple_Adaptor::evalMetric nanoflann.hpp:432-440,
 *                                  RadiusResultSet::addPoint strict '<' :249-253, sort by distance :207-213)
 *
 * Two behaviours of the reference are implementation-defined and are restated from the libraries it uses:
 *   (1) output ORDER of grid_subsampling = iteration order of libstdc++ std::unordered_map<size_t,...>
 *       (identity hash, bucket = key % n_buckets, singly linked node list; GCC 13 bits/hashtable.h
 *       _M_insert_bucket_begin / _M_rehash_aux(unique)).  Restated here as an explicit node-list simulation.
 *       Bucket counts follow std::__detail::_Prime_rehash_policy (max load 1.0, growth 2, sparse prime list):
 *       a rehash to kBuckets[p] happens just before element number kBuckets[p-1]+1 is inserted.
 *   (2) order inside exact-distance tie groups of radius_neighbors (unstable std::sort over KD-tree traversal
 *       order): NOT reproducible; this restatement breaks ties by ascending index.  Ties do not occur on the
 *       continuous synthetic coordinates every test uses (self match d=0 is unique).
 *
 * Pinned against the real reference (oracle/_ref/libref_ext.so) by tests/test_oracle_collate.py.
 * Build: see oracle/Makefile (-ffp-contract=off: every fp32 operation rounds on its own, as on the x86-64
 * build of the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint64_t kBuckets[] = {13ull, 29ull, 59ull, 127ull, 257ull, 541ull, 1109ull, 2357ull, 5087ull,
    10273ull, 20753ull, 42043ull, 85229ull, 172933ull, 351061ull, 712697ull, 1447153ull, 2938679ull, 5967347ull,
    12117689ull, 24607243ull, 49969847ull, 101473717ull, 206062531ull, 418451333ull, 849749479ull, 1725587117ull};
#define N_BUCKET_STEPS ((int)(sizeof(kBuckets) / sizeof(kBuckets[0])))

typedef struct { uint64_t key; int32_t next; int32_t count; float sx, sy, sz; } node_t;

/* node list with a "before begin" pseudo node encoded as index -1; bucket[b] holds the index of the node
 * BEFORE the first node of bucket b (or -2 when empty), exactly like libstdc++'s _M_buckets. */
typedef struct { node_t* nodes; int32_t n_nodes; int32_t head; int32_t* bucket; uint64_t nb; } table_t;

static void place(table_t* t, int32_t id) {
    uint64_t b = t->nodes[id].key % t->nb;
    if (t->bucket[b] != -2) {                       /* bucket in use: insert at the front of its group */
        int32_t before = t->bucket[b];
        int32_t* link = (before == -1) ? &t->head : &t->nodes[before].next;
        t->nodes[id].next = *link;
        *link = id;
    } else {                                        /* new bucket: insert at the front of the whole list */
        t->nodes[id].next = t->head;
        t->head = id;
        if (t->nodes[id].next >= 0) {
            uint64_t nb2 = t->nodes[t->nodes[id].next].key % t->nb;
            t->bucket[nb2] = id;                    /* old first node's bucket now starts after `id` */
        }
        t->bucket[b] = -1;
    }
}

static void rehash(table_t* t, uint64_t nb) {
    free(t->bucket);
    t->bucket = (int32_t*)malloc(sizeof(int32_t) * nb);
    for (uint64_t i = 0; i < nb; ++i) t->bucket[i] = -2;
    t->nb = nb;
    int32_t p = t->head;
    t->head = -1;
    while (p >= 0) {                                /* re-insert in current list order */
        int32_t nx = t->nodes[p].next;
        place(t, p);
        p = nx;
    }
}

static int32_t find(const table_t* t, uint64_t key) {
    if (t->nb == 0) return -1;
    uint64_t b = key % t->nb;
    if (t->bucket[b] == -2) return -1;
    int32_t p = (t->bucket[b] == -1) ? t->head : t->nodes[t->bucket[b]].next;
    while (p >= 0 && t->nodes[p].key % t->nb == b) {
        if (t->nodes[p].key == key) return p;
        p = t->nodes[p].next;
    }
    return -1;
}

static int64_t single_grid_subsampling(const float* pts, int64_t n, float voxel, float* out) {
    /* cloud.cpp:4-37 */
    float mnx = pts[0], mny = pts[1], mnz = pts[2], mxx = pts[0], mxy = pts[1], mxz = pts[2];
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pts + 3 * i;
        if (p[0] < mnx) mnx = p[0];
        if (p[1] < mny) mny = p[1];
        if (p[2] < mnz) mnz = p[2];
        if (p[0] > mxx) mxx = p[0];
        if (p[1] > mxy) mxy = p[1];
        if (p[2] > mxz) mxz = p[2];
    }
    /* grid_subsampling_cpu.cpp:11 : floor(minCorner * (1. / voxel)) * voxel ; the double reciprocal is cast to
     * float by operator*(PointXYZ, float) (cloud.h:92-94). */
    float inv = (float)(1.0 / (double)voxel);
    float ox = floorf(mnx * inv) * voxel, oy = floorf(mny * inv) * voxel, oz = floorf(mnz * inv) * voxel;
    uint64_t NX = (uint64_t)(floorf((mxx - ox) / voxel) + 1.0f);   /* :13-20 */
    uint64_t NY = (uint64_t)(floorf((mxy - oy) / voxel) + 1.0f);
    (void)mxz;

    table_t t;
    t.nodes = (node_t*)malloc(sizeof(node_t) * (size_t)n);
    t.n_nodes = 0; t.head = -1; t.bucket = NULL; t.nb = 0;
    int step = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pts + 3 * i;
        uint64_t ix = (uint64_t)floorf((p[0] - ox) / voxel);        /* :32-34 */
        uint64_t iy = (uint64_t)floorf((p[1] - oy) / voxel);
        uint64_t iz = (uint64_t)floorf((p[2] - oz) / voxel);
        uint64_t key = ix + NX * iy + NX * NY * iz;                 /* :35 */
        int32_t id = find(&t, key);
        if (id < 0) {
            /* _Prime_rehash_policy: grow before the insert that would exceed load factor 1 */
            if (step < N_BUCKET_STEPS && (uint64_t)t.n_nodes == (step == 0 ? 0ull : kBuckets[step - 1])) {
                rehash(&t, kBuckets[step]);
                step++;
            }
            id = t.n_nodes++;
            t.nodes[id].key = key; t.nodes[id].count = 0;
            t.nodes[id].sx = t.nodes[id].sy = t.nodes[id].sz = 0.0f;
            place(&t, id);
        }
        t.nodes[id].count += 1;                                     /* grid_subsampling_cpu.h:17-20 */
        t.nodes[id].sx += p[0]; t.nodes[id].sy += p[1]; t.nodes[id].sz += p[2];
    }
    int64_t m = 0;
    for (int32_t p = t.head; p >= 0; p = t.nodes[p].next) {         /* :44-47, map iteration order */
        float w = (float)(1.0 / (double)t.nodes[p].count);
        out[3 * m + 0] = t.nodes[p].sx * w;
        out[3 * m + 1] = t.nodes[p].sy * w;
        out[3 * m + 2] = t.nodes[p].sz * w;
        ++m;
    }
    free(t.nodes); free(t.bucket);
    return m;
}

/* s_points must hold n_points*3 floats. Returns total subsampled points. */
int64_t oracle_grid_subsampling(const float* points, int64_t n_points, const int64_t* lengths, int64_t batch,
                                float voxel, float* s_points, int64_t* s_lengths) {
    int64_t start = 0, total = 0;
    (void)n_points;
    for (int64_t b = 0; b < batch; ++b) {                           /* grid_subsampling_cpu.cpp:58-72 */
        int64_t m = single_grid_subsampling(points + 3 * start, lengths[b], voxel, s_points + 3 * total);
        s_lengths[b] = m;
        total += m;
        start += lengths[b];
    }
    return total;
}

typedef struct { float d; int64_t i; } cand_t;
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* Brute-force restatement.  Pass out == NULL to obtain the row width (global max neighbour count). */
int64_t oracle_radius_neighbors(const float* q, int64_t nq, const float* s, int64_t ns, const int64_t* ql,
                                const int64_t* sl, int64_t batch, float radius, int64_t* out, int64_t out_cols) {
    float r2 = radius * radius;                                     /* radius_neighbors_cpu.cpp:12 */
    int64_t max_count = 0;
    int64_t qs = 0, ss = 0;
    int64_t max_sl = 0;
    for (int64_t b = 0; b < batch; ++b) if (sl[b] > max_sl) max_sl = sl[b];
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)(max_sl > 0 ? max_sl : 1));
    for (int64_t b = 0; b < batch; ++b) {
        for (int64_t i = qs; i < qs + ql[b]; ++i) {
            int64_t c = 0;
            for (int64_t j = 0; j < sl[b]; ++j) {
                const float* sp = s + 3 * (ss + j);
                float dx = q[3 * i] - sp[0], dy = q[3 * i + 1] - sp[1], dz = q[3 * i + 2] - sp[2];
                /* nanoflann.hpp:432-440: result accumulates diff*diff dimension by dimension, from 0 */
                float d = 0.0f; d += dx * dx; d += dy * dy; d += dz * dz;
                if (d < r2) { cand[c].d = d; cand[c].i = j; ++c; }  /* strict, nanoflann.hpp:249-253 */
            }
            if (c > max_count) max_count = c;
            if (out != NULL) {
                qsort(cand, (size_t)c, sizeof(cand_t), cmp_cand);
                for (int64_t k = 0; k < out_cols; ++k)
                    out[i * out_cols + k] = (k < c) ? cand[k].i + ss : ns;   /* :78-88 offset + sentinel */
            }
        }
        qs += ql[b]; ss += sl[b];
    }
    free(cand);
    (void)nq;
    return max_count;
}
