/*
 * oracle/hashgrid_oracle.c  --  TEST INFRASTRUCTURE ONLY (never linked into / imported by the product).
 *
 * Plain-C CPU restatement (pthreads over independent samples / levels: this image's gcc has no libgomp) of the reference's multi-resolution
 * hash-grid encoder, which the reference only ships as CUDA (no CPU path exists in
 * im2scene/sdf/models/gridencoder/grid.py).  Every function cites the reference file:line it follows;
 * paths are relative to /root/reference/im2scene/sdf/models/gridencoder/.
 *
 * Parity status: the reference holds NO golden vectors for this path, so this restatement is
 * pinned on the GPU box against the UNMODIFIED reference kernels compiled into oracle/_ref/_gridencoder_ref.so
 * (tests/test_gpu_parity_ref.py); on CPU it is "parity unpinned" by construction.
 *
 * Arithmetic notes that matter for bit-exact indices (SURVEY.md section 7 "hard parts"):
 *   - scale = exp2f(level*S)*H - 1 is computed ON DEVICE by the reference (src/gridencoder.cu:138) with CUDA's
 *     exp2f, whose last bit may differ from libm.  Callers may therefore pass the device-computed per-level scale
 *     table in `level_scales`; if NULL, libm exp2f is used.
 *   - pos = x*scale + 0.5 is FMA-contracted by nvcc (-fmad=true default) -> fmaf here (src/gridencoder.cu:148).
 *   - the weighted sum `results += w * grid[...]` is likewise contracted -> fmaf (src/gridencoder.cu:187).
 *   - index arithmetic is uint32 with wraparound (src/gridencoder.cu:50-84).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORACLE_MAX_D 3
#define ORACLE_MAX_C 8

/* Minimal parallel-for: [0, n) in contiguous chunks over min(threads, n / grain) pthreads.  ORACLE_THREADS overrides the count
 * (default: online cores, at most 64).  The work functions write disjoint outputs, so results do not depend on the thread count. */
typedef void (*range_fn)(int64_t begin, int64_t end, void* ctx);
typedef struct { range_fn fn; void* ctx; int64_t begin, end; } pf_task;
static void* pf_run(void* p) { pf_task* t = (pf_task*)p; t->fn(t->begin, t->end, t->ctx); return NULL; }
int oracle_num_threads(void) {
    const char* e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return (int)(n < 1 ? 1 : (n > 64 ? 64 : n));
}
static void parallel_for(int64_t n, int64_t grain, range_fn fn, void* ctx) {
    int64_t T = oracle_num_threads();
    if (grain < 1) grain = 1;
    if (T > (n + grain - 1) / grain) T = (n + grain - 1) / grain;
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t th[64]; pf_task task[64];
    const int64_t per = (n + T - 1) / T;
    int64_t started = 0;
    for (int64_t t = 0; t < T; t++) {
        task[t].fn = fn; task[t].ctx = ctx; task[t].begin = t * per; task[t].end = (t + 1) * per < n ? (t + 1) * per : n;
        if (task[t].begin >= task[t].end) break;
        if (pthread_create(&th[t], NULL, pf_run, &task[t]) != 0) { fn(task[t].begin, n, ctx); break; }   /* fall back to this thread */
        started++;
    }
    for (int64_t t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* src/gridencoder.cu:50-63 : coherent prime hash, uint32 wraparound */
static uint32_t fast_hash(uint32_t D, const uint32_t* pos_grid) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= pos_grid[i] * primes[i];
    return r;
}

/* src/gridencoder.cu:66-84 : dense stride walk while stride <= hashmap_size, else hash; modulo; times C */
static uint32_t grid_index(uint32_t D, uint32_t C, uint32_t gridtype, int align_corners, uint32_t ch,
                           uint32_t hashmap_size, uint32_t resolution, const uint32_t* pos_grid) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(D, pos_grid);
    return (index % hashmap_size) * C + ch;
}

/* src/gridencoder.cu:138-139 */
static void level_geometry(uint32_t level, float S, uint32_t H, const float* level_scales, float* scale, uint32_t* resolution) {
    float sc = level_scales ? level_scales[level] : (exp2f((float)level * S) * (float)H - 1.0f);
    *scale = sc;
    *resolution = (uint32_t)ceilf(sc) + 1;
}

/* Write the per-level scale table exactly as the restatement computes it (libm). */
void oracle_grid_level_scales(uint32_t L, float S, uint32_t H, float* out) {
    for (uint32_t l = 0; l < L; l++) out[l] = exp2f((float)l * S) * (float)H - 1.0f;
}

static float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }           /* src/gridencoder.cu:39-42 */
static float smoothstep_d(float v) { return 6 * v * (1.0f - v); }                  /* src/gridencoder.cu:44-47 */

/* shared prologue: src/gridencoder.cu:110-159.  returns 1 if out of bounds */
static int locate(uint32_t D, const float* x, float scale, int align_corners, uint32_t interp,
                  float* pos, float* pos_deriv, uint32_t* pos_grid) {
    for (uint32_t d = 0; d < D; d++)
        if (x[d] < 0 || x[d] > 1) return 1;
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
        pos_grid[d] = (uint32_t)floorf(pos[d]);
        pos[d] -= (float)pos_grid[d];
        if (interp == 1) { pos_deriv[d] = smoothstep_d(pos[d]); pos[d] = smoothstep_f(pos[d]); }
        else pos_deriv[d] = 1.0f;
    }
    return 0;
}

/*
 * Forward: src/gridencoder.cu:87-245 (kernel_grid) + launcher :372-383.
 *   inputs      [B, D] in [0,1]
 *   embeddings  [sum(offsets), C]
 *   offsets     [L+1]
 *   outputs     [L, B, C]   (the reference's L-major layout, grid.py:47)
 *   dy_dx       [B, L, D, C] or NULL
 *   corner_idx  optional [B, L, 2^D] uint32 : table ROW index (without the *C) of each corner, 0xFFFFFFFF if OOB
 *   corner_w    optional [B, L, 2^D] float  : D-linear weight of each corner
 */
static void grid_encode_forward_range(int64_t b_begin, int64_t b_end, const float* inputs, const float* embeddings, const int* offsets,
                                      float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      float* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp,
                                      const float* level_scales, uint32_t* corner_idx, float* corner_w) {
    const uint32_t NC = 1u << D;
    /* samples are independent: split over the host threads by the caller (fwd_range / parallel_for); results do not depend on the count */
    for (int64_t bb = b_begin; bb < b_end; bb++) {
        const uint32_t b = (uint32_t)bb;
        const float* x = inputs + (size_t)b * D;
        for (uint32_t level = 0; level < L; level++) {
            const float* grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
            float* out = outputs + ((size_t)level * B + b) * C;
            float* dd = dy_dx ? dy_dx + ((size_t)b * L + level) * D * C : NULL;
            uint32_t* ci = corner_idx ? corner_idx + ((size_t)b * L + level) * NC : NULL;
            float* cw = corner_w ? corner_w + ((size_t)b * L + level) * NC : NULL;
            const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
            float scale; uint32_t resolution;
            level_geometry(level, S, H, level_scales, &scale, &resolution);
            float pos[ORACLE_MAX_D], pos_deriv[ORACLE_MAX_D]; uint32_t pos_grid[ORACLE_MAX_D];
            if (locate(D, x, scale, align_corners, interp, pos, pos_deriv, pos_grid)) {   /* :118-135 */
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = 0;
                if (ci) for (uint32_t i = 0; i < NC; i++) { ci[i] = 0xFFFFFFFFu; cw[i] = 0.0f; }
                continue;
            }
            float results[ORACLE_MAX_C] = {0};
            for (uint32_t idx = 0; idx < NC; idx++) {                                     /* :166-191 */
                float w = 1; uint32_t pgl[ORACLE_MAX_D];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                uint32_t index = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                for (uint32_t ch = 0; ch < C; ch++) results[ch] = fmaf(w, grid[index + ch], results[ch]);
                if (ci) { ci[idx] = index / C; cw[idx] = w; }
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = results[ch];
            if (dd) {                                                                     /* :201-244 */
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[ORACLE_MAX_C] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = scale; uint32_t pgl[ORACLE_MAX_D];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                            else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                        }
                        pgl[gd] = pos_grid[gd];
                        uint32_t il = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                        pgl[gd] = pos_grid[gd] + 1;
                        uint32_t ir = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                        for (uint32_t ch = 0; ch < C; ch++)
                            rg[ch] = fmaf(w * (grid[ir + ch] - grid[il + ch]), pos_deriv[gd], rg[ch]);
                    }
                    for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = rg[ch];
                }
            }
        }
    }
}

typedef struct {
    const float* inputs; const float* embeddings; const int* offsets; float* outputs; uint32_t B, D, C, L; float S; uint32_t H;
    float* dy_dx; uint32_t gridtype; int align_corners; uint32_t interp; const float* level_scales; uint32_t* corner_idx; float* corner_w;
} fwd_args;
static void fwd_range(int64_t a, int64_t b, void* p) {
    fwd_args* q = (fwd_args*)p;
    grid_encode_forward_range(a, b, q->inputs, q->embeddings, q->offsets, q->outputs, q->B, q->D, q->C, q->L, q->S, q->H, q->dy_dx, q->gridtype,
                              q->align_corners, q->interp, q->level_scales, q->corner_idx, q->corner_w);
}
void oracle_grid_encode_forward(const float* inputs, const float* embeddings, const int* offsets, float* outputs,
                                uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                float* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp,
                                const float* level_scales, uint32_t* corner_idx, float* corner_w) {
    fwd_args a = {inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp, level_scales, corner_idx, corner_w};
    parallel_for((int64_t)B, 4096, fwd_range, &a);
}

/*
 * Backward: src/gridencoder.cu:248-340 (kernel_grid_backward) + :343-369 (kernel_input_backward).
 *   grad            [L, B, C]
 *   grad_embeddings [sum(offsets), C]  accumulated INTO (caller pre-zeroes, grid.py:77)
 *   grad_inputs     [B, D] or NULL     overwritten (needs dy_dx)
 * The GPU reference uses float atomics whose order is non-deterministic; this restatement accumulates in double and
 * rounds once, i.e. it is the exact sum the atomics approximate (tolerance-compared, never bit-compared).
 */
typedef struct {
    const float* grad; const float* inputs; const int* offsets; double* acc; uint32_t B, D, C, L; float S; uint32_t H;
    uint32_t gridtype; int align_corners; uint32_t interp; const float* level_scales; const float* dy_dx; float* grad_inputs;
} bwd_args;

/* one table level = one disjoint region of the accumulator: the serial accumulation order inside a level is kept */
static void bwd_levels(int64_t l_begin, int64_t l_end, void* p) {
    const bwd_args* q = (const bwd_args*)p;
    const uint32_t B = q->B, D = q->D, C = q->C, NC = 1u << q->D;
    for (int64_t lv = l_begin; lv < l_end; lv++) {
        const uint32_t level = (uint32_t)lv;
        double* gg = q->acc + (size_t)(uint32_t)q->offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(q->offsets[level + 1] - q->offsets[level]);
        float scale; uint32_t resolution;
        level_geometry(level, q->S, q->H, q->level_scales, &scale, &resolution);
        for (uint32_t b = 0; b < B; b++) {
            const float* x = q->inputs + (size_t)b * D;
            const float* g = q->grad + ((size_t)level * B + b) * C;
            float pos[ORACLE_MAX_D], pos_deriv[ORACLE_MAX_D]; uint32_t pos_grid[ORACLE_MAX_D];
            if (locate(D, x, scale, q->align_corners, q->interp, pos, pos_deriv, pos_grid)) continue;   /* :276-281 */
            for (uint32_t idx = 0; idx < NC; idx++) {                                             /* :305-339 */
                float w = 1; uint32_t pgl[ORACLE_MAX_D];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                uint32_t index = grid_index(D, C, q->gridtype, q->align_corners, 0, hashmap_size, resolution, pgl);
                for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += (double)(w * g[ch]);
            }
        }
    }
}

static void bwd_inputs(int64_t b_begin, int64_t b_end, void* p) {                               /* :343-369 */
    const bwd_args* q = (const bwd_args*)p;
    const uint32_t B = q->B, D = q->D, C = q->C, L = q->L;
    for (int64_t b = b_begin; b < b_end; b++)
        for (uint32_t d = 0; d < D; d++) {
            float r = 0;
            for (uint32_t l = 0; l < L; l++)
                for (uint32_t ch = 0; ch < C; ch++)
                    r = fmaf(q->grad[((size_t)l * B + (size_t)b) * C + ch], q->dy_dx[(((size_t)b * L + l) * D + d) * C + ch], r);
            q->grad_inputs[(size_t)b * D + d] = r;
        }
}

void oracle_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int* offsets,
                                 float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                 const float* dy_dx, float* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                 const float* level_scales) {
    (void)embeddings;
    const size_t total = (size_t)(uint32_t)offsets[L] * C;
    double* acc = (double*)calloc(total, sizeof(double));
    bwd_args a = {grad, inputs, offsets, acc, B, D, C, L, S, H, gridtype, align_corners, interp, level_scales, dy_dx, grad_inputs};
    parallel_for((int64_t)L, 1, bwd_levels, &a);
    for (size_t i = 0; i < total; i++) grad_embeddings[i] += (float)acc[i];
    free(acc);
    if (dy_dx && grad_inputs) parallel_for((int64_t)B, 4096, bwd_inputs, &a);
}

/*
 * Total-variation gradient: src/gridencoder.cu:506-610 (kernel_grad_tv).  Adds into `grad` [sum(offsets), C].
 * Accumulated in double for the same reason as above.
 */
void oracle_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int* offsets, float weight,
                                 uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                 uint32_t gridtype, int align_corners, const float* level_scales) {
    const size_t total = (size_t)(uint32_t)offsets[L] * C;
    double* acc = (double*)calloc(total, sizeof(double));
    for (uint32_t level = 0; level < L; level++) {
        const float* grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        double* gg = acc + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        float scale; uint32_t resolution;
        level_geometry(level, S, H, level_scales, &scale, &resolution);
        for (uint32_t b = 0; b < B; b++) {
            const float* x = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            uint32_t pos_grid[ORACLE_MAX_D];
            for (uint32_t d = 0; d < D; d++)
                pos_grid[d] = (uint32_t)floorf(fmaf(x[d], scale, align_corners ? 0.0f : 0.5f));   /* :548-553 */
            float results[ORACLE_MAX_C] = {0}, idelta[ORACLE_MAX_C] = {0};
            uint32_t index = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
            float w = weight / (2 * D);
            for (uint32_t d = 0; d < D; d++) {                                                    /* :565-601 */
                uint32_t cur = pos_grid[d];
                if (cur < resolution) {
                    pos_grid[d] = cur + 1;
                    uint32_t ir = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        float gv = grid[index + ch] - grid[ir + ch];
                        results[ch] += gv; idelta[ch] = fmaf(gv, gv, idelta[ch]);
                    }
                }
                if (cur > 0) {
                    pos_grid[d] = cur - 1;
                    uint32_t il = grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        float gv = grid[index + ch] - grid[il + ch];
                        results[ch] += gv; idelta[ch] = fmaf(gv, gv, idelta[ch]);
                    }
                }
                pos_grid[d] = cur;
            }
            for (uint32_t ch = 0; ch < C; ch++)
                gg[index + ch] += (double)(w * results[ch] * (1.0f / sqrtf(idelta[ch] + 1e-9f)));  /* :607 */
        }
    }
    for (size_t i = 0; i < total; i++) grad[i] += (float)acc[i];
    free(acc);
}
