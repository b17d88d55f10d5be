// pool_kernel.cuh -- the wavefront pool kernel template.  Included by kernels.cu (ahead-of-time
// variants) and by the scene-specialised translation units that jit.cu compiles with NVRTC.
#pragma once
#include "kernels_config.h"
#include "rt_integrator.cuh"

namespace rt {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// Wavefront pool kernel.
//
// Every warp owns a pool of NSLOT path slots in shared memory (structure of arrays, one word
// per field per slot).  Families A/B: a slot traces ONE sample (work item = (pixel, sample)),
// writes its radiance to the per-sample scratch buffer in HBM and pulls the next item from the
// global work queue; k_fold_samples then adds a pixel's samples in sample order, so the fp32
// accumulation order equals the reference's launch-by-launch `buffer += color` while the work
// granularity is a single path (no ragged tail: 67 M items on C1).  Family C: a slot is a pixel
// worker, because its launches are sequentially dependent through ray_buffer.
//
// The warp alternates between phases, all at (nearly) full lane occupancy:
//   MARCH    each lane holds the march state of one slot in registers and the warp-wide loop
//            body is ONE sphere-tracing step (scene SDF evaluation) followed by one vote: "did
//            any lane's march end?".  Nothing else lives in the loop -- how a march ended is
//            decided afterwards (march_status), lanes without a slot march a harmless dummy ray
//            and are masked out of the vote (idle_march), constants are literals in the
//            specialised kernels.  (Bunny scenes: two stages per iteration, see pool_body.)
//   FINISH   the lanes whose march ended write t_eval + status to their slots, push them on the
//   + REFILL warp's `pending` stack and pop ready-to-march slots from the `ready` stack right
//            there (ballot + popc ranks, no atomics); the loop is re-entered directly.
//   RESOLVE  when lanes would idle (ready stack empty) the marching lanes park their slots and
//            all 32 lanes pop pending slots (resolve_batch): surface interaction (normal, BSDF
//            sample, Russian roulette), sample output, path regeneration, and pulls from the
//            warp's chunk of the work queue.  Resolved slots go back on the ready stack.
// With NSLOT = 64 the pool holds 32 marching + 32 ready/pending slots, so the heavy-tailed
// march lengths (p50 28, p99 ~100 steps) no longer idle lanes, and the divergent shading code
// runs on compacted batches.  The RNG is keyed by (pixel, sample, draw index) only, so the
// scheduling cannot change any number a sample sees: results are bit-identical to the simple
// kernel and to the CPU oracle.
//
// Round 2 (scene-specialised builds, switched on by the code generator / capi.cu: jit_build_options):
//   RT_JIT_FAST   the march loop evaluates jit_nearest_fast() (walls as planes inside a proven region).  A lane whose
//                 point lies outside the region "finishes" with status ST_SLOW, its step taken back; the full-code
//                 steps it needs are taken by slow_march() in the resolve phase.
//   RT_JIT_TSTOP  per-ray t_stop (slot word F_TSTOP): beyond it the ray provably misses (ray_t_stop).
//   RT_REGEN_MIN  RESOLVE splits into two batch types.  MODE_HITS batches shade hits / misses and start the next bounce:
//                 no work-queue pulls, no new paths, no full-code march steps.  MODE_FRESH ("regeneration") batches
//                 take the slots of the third stack, `fresh`: ended paths (fetch + camera ray + first roulette + the
//                 camera ray's full-code steps towards the region), ST_SLOW drop-outs and the rays that missed (their
//                 path ends: sky colour, sample written, regeneration -- nothing of a hit batch).  They run when RT_REGEN_MIN
//                 slots wait or RT_REGEN_IDLE lanes have nothing to march.
//   RT_FIN_MIN    finished lanes wait (masked) until RT_FIN_MIN of them can leave the march loop together.
// ------------------------------------------------------------------------------------------
enum : int { ST_NONE = 0, ST_READY = 1, ST_HIT = 2, ST_MISS = 3, ST_DONE = 4, ST_FETCH = 5, ST_NEWPATH = 6,
             ST_ADVANCE = 7, ST_DEAD = 8, ST_SLOW = 9 };

// slot fields (word index into the per-warp SoA)
enum : int { F_ROX = 0, F_ROY, F_ROZ, F_RDX, F_RDY, F_RDZ, F_COLX, F_COLY, F_COLZ, F_T, F_W, F_S, F_D, F_TEVAL,
             F_STEPS, F_IDX, F_DEPTH, F_RNGN, F_PIXEL, F_SAMP, F_K, F_STATUS, F_ACCX, F_ACCY, F_ACCZ, F_ACCW,
             F_COUNT };
static_assert(F_COUNT == kPoolSlotWords, "kernels_config.h: kPoolSlotWords");
constexpr int F_TSTOP = F_IDX;   // families A/B (the argmin is re-evaluated at the hit): the ray's t_stop, see ray_t_stop()
// per-warp work-queue chunk (words after the two stacks)
enum : int { WQ_LO = 0, WQ_HI, WQ_LEFT, WQ_ITEM, WQ_SAMP, WQ_NFRESH, WQ_COUNT };
static_assert(WQ_COUNT <= kPoolQueueWords, "kernels_config.h: kPoolQueueWords");

template <int NSLOT>
struct Pool {
    uint32_t* w;   // [F_COUNT][NSLOT]
    __device__ __forceinline__ float getf(int f, int slot) const { return __uint_as_float((w + slot)[f * NSLOT]); }
    __device__ __forceinline__ int geti(int f, int slot) const { return (int)(w + slot)[f * NSLOT]; }
    __device__ __forceinline__ void setf(int f, int slot, float v) { (w + slot)[f * NSLOT] = __float_as_uint(v); }
    __device__ __forceinline__ void seti(int f, int slot, int v) { (w + slot)[f * NSLOT] = (uint32_t)v; }
};

// A slot's march state is split by who needs it.  Marching needs the ray, t, the step count and the
// relaxation state; t_eval and idx are rewritten by every march step before anybody reads them.  The resolve
// phase needs the ray and, families A/B, t_eval (hit position; the argmin is re-evaluated there), family C
// the marched origin and the argmin.  The step count only feeds the work counters there.
template <class VAR, int NSLOT>
__device__ __forceinline__ void load_ray_of(const Pool<NSLOT>& pool, int slot, MarchState& m)
{
    m.ro = V3(pool.getf(F_ROX, slot), pool.getf(F_ROY, slot), pool.getf(F_ROZ, slot));
    m.rd = V3(pool.getf(F_RDX, slot), pool.getf(F_RDY, slot), pool.getf(F_RDZ, slot));
}
template <class VAR, int NSLOT>
__device__ __forceinline__ void load_march(const Pool<NSLOT>& pool, int slot, MarchState& m)
{
    load_ray_of<VAR, NSLOT>(pool, slot, m);
    m.t = pool.getf(F_T, slot);
    m.steps = pool.geti(F_STEPS, slot);
    if (VAR::MARCHER != MARCH_PLAIN) {
        m.w = pool.getf(F_W, slot); m.s = pool.getf(F_S, slot); m.d = pool.getf(F_D, slot);
    }
#if defined(RT_JIT_TSTOP)
    if (VAR::MARCHER != MARCH_SRC) m.t_stop = pool.getf(F_TSTOP, slot);
#endif
}
template <class VAR, int NSLOT>
__device__ __forceinline__ void load_finished(const Pool<NSLOT>& pool, int slot, MarchState& m)
{
    load_ray_of<VAR, NSLOT>(pool, slot, m);
    if (VAR::MARCHER == MARCH_SRC) m.idx = pool.geti(F_IDX, slot);
    else m.t_eval = pool.getf(F_TEVAL, slot);
    if (VAR::COUNT) m.steps = pool.geti(F_STEPS, slot);
}
// a marching lane parks its slot (the ray itself is already in the slot, except family C's marched origin)
template <class VAR, int NSLOT>
__device__ __forceinline__ void store_parked(Pool<NSLOT>& pool, int slot, const MarchState& m)
{
    if (VAR::MARCHER == MARCH_SRC) {
        pool.setf(F_ROX, slot, m.ro.x); pool.setf(F_ROY, slot, m.ro.y); pool.setf(F_ROZ, slot, m.ro.z);
    }
    pool.setf(F_T, slot, m.t);
    pool.seti(F_STEPS, slot, m.steps);
    if (VAR::MARCHER != MARCH_PLAIN) {
        pool.setf(F_W, slot, m.w); pool.setf(F_S, slot, m.s); pool.setf(F_D, slot, m.d);
    }
}
// a freshly begun bounce (march_begin() state) goes back to the pool
template <class VAR, int NSLOT>
__device__ __forceinline__ void store_ready(Pool<NSLOT>& pool, int slot, const MarchState& m)
{
    pool.setf(F_ROX, slot, m.ro.x); pool.setf(F_ROY, slot, m.ro.y); pool.setf(F_ROZ, slot, m.ro.z);
    pool.setf(F_RDX, slot, m.rd.x); pool.setf(F_RDY, slot, m.rd.y); pool.setf(F_RDZ, slot, m.rd.z);
    pool.setf(F_T, slot, m.t);
    pool.seti(F_STEPS, slot, m.steps);
#if defined(RT_JIT_TSTOP)
    if (VAR::MARCHER != MARCH_SRC) pool.setf(F_TSTOP, slot, m.t_stop);
#endif
    if (VAR::MARCHER != MARCH_PLAIN) {
        pool.setf(F_W, slot, m.w); pool.setf(F_S, slot, m.s); pool.setf(F_D, slot, m.d);
    }
}
template <class VAR, int NSLOT>
__device__ __forceinline__ void store_finished(Pool<NSLOT>& pool, int slot, const MarchState& m, int st)
{
    if (VAR::MARCHER == MARCH_SRC) {
        pool.setf(F_ROX, slot, m.ro.x); pool.setf(F_ROY, slot, m.ro.y); pool.setf(F_ROZ, slot, m.ro.z);
        pool.seti(F_IDX, slot, m.idx);
    } else {
        pool.setf(F_TEVAL, slot, m.t_eval);
    }
    if (VAR::COUNT) pool.seti(F_STEPS, slot, m.steps);
    pool.seti(F_STATUS, slot, st);
}

// March state of a lane that holds no slot: a ray that stays at (8, 8, 8) -- outside every primitive's
// expensive region (the neural bunny lives in the unit sphere) -- so the march loop needs no branch around
// idle lanes; their results are masked out of the vote.  (t stays finite for ~1e36 steps; nothing else of
// an idle lane's state is ever read.)
__device__ __forceinline__ void idle_march(MarchState& m)
{
    m.ro = V3(8.0f); m.rd = V3(0.0f);
}
__device__ __forceinline__ void zero_march(MarchState& m)
{
    idle_march(m);
    m.t = 0.0f; m.w = 1.0f; m.s = 0.0f; m.d = 0.0f; m.t_eval = 0.0f;
    m.steps = 0; m.idx = 0; m.t_stop = 3.0e38f;
}

// One resolve batch: every lane with slot >= 0 runs its slot's state machine (surface interaction, sample
// accumulation, path regeneration, work-queue pull) until the slot needs marching again or dies; slots that
// are ready to march go on the ready stack.
// RT_REGEN_MIN > 0 (experiment, families A/B; NVRTC builds: env RTPBR_REGEN_MIN): path regeneration leaves the
// hit / miss batches.  A slot whose path ended goes on a third stack (`fresh`) and regeneration batches run only when
// RT_REGEN_MIN slots wait there or lanes would idle, so that fetch + camera ray + first roulette run at ~25 instead
// of ~8 lanes.  MODE_HITS batches never fetch, MODE_FRESH batches only fetch and start paths.
#ifndef RT_REGEN_MIN
#define RT_REGEN_MIN 0
#endif
enum : int { MODE_ALL = 0, MODE_HITS = 1, MODE_FRESH = 2 };
#ifndef RT_FIN_MIN
#define RT_FIN_MIN 1
#endif
#ifndef RT_SLOW_FIRST
#define RT_SLOW_FIRST 1      // regeneration batches: full-code march steps at the top of the state machine (0: at its end, as before)
#endif
#ifndef RT_REGEN_IDLE
#define RT_REGEN_IDLE 1      // regeneration batches also run when at least this many lanes have nothing to march
#endif

template <class VAR, int NSLOT, int MODE = MODE_ALL>
__device__ __forceinline__ void resolve_batch(const KParams& P, Pool<NSLOT>& pool, const int slot, const int lane,
                                              const unsigned lane_lt, uint8_t* ready, int& n_ready, volatile uint32_t* wq,
                                              WorkCounters& cnt, uint8_t* pend = nullptr, int* n_pend = nullptr, uint8_t* fresh = nullptr,
                                              int* n_fresh = nullptr)
{
    // ---- load the slot
    Path p;
    int st = ST_NONE, samp = 0, k = 0;
    uint32_t pixel = 0;
    unsigned long long wid = 0;                       // families A/B: work item = scratch index
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);     // family C: the pixel's image_buffer entry
    zero_march(p.m);
    p.col = V3(0.f);
    p.depth = 0;
    p.rng = rng_make(0u, 0u, 0u);
    if (slot >= 0) {
        st = pool.geti(F_STATUS, slot);
#if defined(RT_JIT_FAST)
        if (st == ST_SLOW) load_march<VAR, NSLOT>(pool, slot, p.m);   // dropped out of the march loop: the march goes on here
        else
#endif
        load_finished<VAR, NSLOT>(pool, slot, p.m);
        p.col = V3(pool.getf(F_COLX, slot), pool.getf(F_COLY, slot), pool.getf(F_COLZ, slot));
        p.depth = pool.geti(F_DEPTH, slot);
        pixel = (uint32_t)pool.geti(F_PIXEL, slot);
        samp = pool.geti(F_SAMP, slot);
        k = pool.geti(F_K, slot);
        p.rng = rng_make(pixel, P.sample_base + (uint32_t)samp, (uint32_t)pool.geti(F_RNGN, slot));
        if (VAR::FAMILY == FAMILY_C)
            acc = make_float4(pool.getf(F_ACCX, slot), pool.getf(F_ACCY, slot), pool.getf(F_ACCZ, slot),
                              pool.getf(F_ACCW, slot));
        else
            wid = (unsigned long long)(uint32_t)pool.geti(F_ACCX, slot) |
                  ((unsigned long long)(uint32_t)pool.geti(F_ACCY, slot) << 32);
    }
    int pi = (int)(pixel / (uint32_t)RT_CFG(P, height)), pj = (int)(pixel - (uint32_t)pi * (uint32_t)RT_CFG(P, height));

    // ---- run the slot's state machine until it needs marching again (or dies)
    for (;;) {
        if (MODE == MODE_FRESH) {
#if defined(RT_JIT_FAST) && RT_SLOW_FIRST
            // Full-code march steps FIRST: the drop-outs of the march loop and the primary rays begun in the previous
            // iteration (below) take them together, and a ray that turns out to have left the scene ends its path and is
            // regenerated in THIS iteration, with everybody else -- not in an iteration of its own at a few lanes.
            if (st == ST_SLOW) {
                const int pre = slow_march<VAR>(P, p.m);
                st = pre == MARCH_CONTINUE ? ST_READY : (pre == MARCH_HIT ? ST_HIT : ST_MISS);
            }
#endif
        }
        const int st_in = st;
        if (MODE == MODE_FRESH) {
            // regeneration batches carry no hits; missed rays end their path here (they come straight from the march loop,
            // or from slow_march()) and the slot fetches a new path in the same pass
            if (VAR::FAMILY != FAMILY_C && st == ST_MISS) {
                if (VAR::COUNT) { cnt.rays++; cnt.evals += (unsigned long long)p.m.steps; }
                on_miss<VAR>(P, p);
                st = ST_DONE;
            }
        } else if (VAR::FAMILY != FAMILY_C) {
            if (st == ST_HIT) {
                if (VAR::COUNT) { cnt.normals++; cnt.rays++; cnt.evals += (unsigned long long)p.m.steps; }
                st = (on_hit<VAR>(P, p) && begin_bounce<VAR>(P, p)) ? ST_READY : ST_DONE;
            } else if (st == ST_MISS) {
                if (VAR::COUNT) { cnt.rays++; cnt.evals += (unsigned long long)p.m.steps; }
                on_miss<VAR>(P, p);
                st = ST_DONE;
            }
        } else if (st == ST_HIT || st == ST_MISS) {
            c_after_march<VAR>(P, p, st == ST_HIT ? MARCH_HIT : MARCH_MISS, VAR::COUNT ? &cnt : nullptr);
            k++;
            st = ST_ADVANCE;
        }
        if (VAR::FAMILY == FAMILY_C) {
            if (st == ST_ADVANCE) {
                TaskC task; task.launch = samp; task.k = k;
                if (c_advance<VAR>(P, pi, pj, p, task, acc, VAR::COUNT ? &cnt : nullptr)) {
                    st = ST_READY;
                    k = task.k;
                } else if (++samp == P.spp) {      // all reference launches replayed: persist the ray
                    store_ray(P.ray_buffer + (size_t)pixel * 10, p);
                    P.image_buffer[pixel] = acc;
                    st = ST_FETCH;
                } else {
                    p.rng = rng_make(pixel, P.sample_base + (uint32_t)samp, 0u);
                    k = 0;
                }
            }
        } else if (st == ST_DONE) {
            P.scratch[wid] = make_float4(p.col.x, p.col.y, p.col.z, 1.0f);   // vec4(ray.color, 1.0)
            st = ST_FETCH;
        }
        // Pull from the work queue (tile padding is skipped).  The warp owns a CHUNK of consecutive work
        // items (wq[]: next item, items left, and for families A/B the (pixel item, sample) pair of the next
        // item, so that the 64-bit division happens once per chunk); the global counter is only touched
        // when the chunk runs out.  Chunk sizes shrink towards the end of the queue (guided scheduling).
        for (;;) {
            if (MODE == MODE_HITS) break;                       // ended paths wait on the fresh stack
            const unsigned m_fetch = __ballot_sync(kFull, st == ST_FETCH);
            if (m_fetch == 0u) break;
            unsigned left = __shfl_sync(kFull, wq[WQ_LEFT], 0);   // (shfl: provably warp-uniform for the compiler)
            if (left == 0u) {                                   // warp-uniform
                const unsigned long long total = VAR::FAMILY == FAMILY_C
                    ? (unsigned long long)P.total_work : (unsigned long long)P.total_work * (unsigned long long)P.spp;
                unsigned long long base = 0;
                unsigned chunk = 0;
                __syncwarp();                                   // every lane has read wq[] before lane 0 rewrites it
                if (lane == 0) {
                    const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(P.work_counter);
                    const unsigned long long rest = total > seen ? total - seen : 0ull;
                    const unsigned long long fair = rest / ((unsigned long long)(gridDim.x * (blockDim.x >> 5)) * 4ull);
                    constexpr unsigned kMax = VAR::FAMILY == FAMILY_C ? 32u : 256u;
                    chunk = fair > (unsigned long long)kMax ? kMax : (fair < 32ull ? 32u : (unsigned)fair);
                    base = atomicAdd(P.work_counter, (unsigned long long)chunk);
                    if (base >= total) chunk = 0u;
                    else if (base + chunk > total) chunk = (unsigned)(total - base);
                    wq[WQ_LO] = (uint32_t)base; wq[WQ_HI] = (uint32_t)(base >> 32); wq[WQ_LEFT] = chunk;
                    if (VAR::FAMILY != FAMILY_C && chunk != 0u) {
                        const unsigned long long item = base / (unsigned long long)P.spp;
                        wq[WQ_ITEM] = (uint32_t)item;
                        wq[WQ_SAMP] = (uint32_t)(base - item * (unsigned long long)P.spp);
                    }
                }
                __syncwarp();
                left = __shfl_sync(kFull, wq[WQ_LEFT], 0);
                if (left == 0u) {                               // queue exhausted
                    if (st == ST_FETCH) st = ST_DEAD;
                    break;
                }
            }
            const unsigned long long cur = (unsigned long long)wq[WQ_LO] | ((unsigned long long)wq[WQ_HI] << 32);
            const uint32_t item0 = wq[WQ_ITEM], samp0 = wq[WQ_SAMP];
            const unsigned want = (unsigned)__popc(m_fetch);
            const unsigned n_take = want < left ? want : left;
            const unsigned r = (unsigned)__popc(m_fetch & lane_lt);
            __syncwarp();
            if (lane == 0) {
                const unsigned long long nxt = cur + n_take;
                wq[WQ_LO] = (uint32_t)nxt; wq[WQ_HI] = (uint32_t)(nxt >> 32); wq[WQ_LEFT] = left - n_take;
                if (VAR::FAMILY != FAMILY_C) {
                    const uint32_t s2 = samp0 + n_take, q2 = s2 / (uint32_t)P.spp;
                    wq[WQ_ITEM] = item0 + q2; wq[WQ_SAMP] = s2 - q2 * (uint32_t)P.spp;
                }
            }
            __syncwarp();
            if (st == ST_FETCH && r < n_take) {
                const unsigned long long wk = cur + r;
                if (VAR::FAMILY == FAMILY_C) {          // work item = pixel
                    if (work_to_pixel(P, (uint32_t)wk, pi, pj) &&
                        // src/pathtracer.py:97-101: if diff > NOISE_THRESHOLD: sample(i, j)
                        (!P.adaptive || P.diff_pixels[pi * RT_CFG(P, height) + pj] > P.noise_threshold)) {
                        pixel = (uint32_t)(pi * RT_CFG(P, height) + pj);
                        acc = P.image_buffer[pixel];
                        samp = 0;
                        k = 0;
                        load_ray(P.ray_buffer + (size_t)pixel * 10, p);
                        p.rng = rng_make(pixel, P.sample_base, 0u);
                        st = ST_ADVANCE;
                    }
                } else {                                // work item = (pixel item, sample)
                    const uint32_t s1 = samp0 + r, q1 = s1 / (uint32_t)P.spp;
                    if (work_to_pixel(P, item0 + q1, pi, pj)) {
                        pixel = (uint32_t)(pi * RT_CFG(P, height) + pj);
                        samp = (int)(s1 - q1 * (uint32_t)P.spp);
                        wid = wk;
                        st = ST_NEWPATH;
                    }
                }
            }
        }
        if (MODE != MODE_HITS && VAR::FAMILY != FAMILY_C && st == ST_NEWPATH) {
            if (VAR::COUNT) cnt.samples++;
            begin_path<VAR>(P, pixel, pi, pj, P.sample_base + (uint32_t)samp, p);
            st = begin_bounce<VAR>(P, p) ? ST_READY : ST_DONE;
        }
#if defined(RT_JIT_SCENE)
        // A bounce has just begun (or, fast-region kernels, a lane dropped out of the march loop: ST_SLOW).
        // Irregular rays (non-finite origin / direction) never enter the specialised march loop; the others get
        // their t_stop (scene bounds) and take the steps that lie outside the fast region right here.
        {
            const bool begun = st == ST_READY && st_in != ST_READY;
            if (begun && ray_is_irregular(p.m)) {
                st = march_to_end_generic<VAR>(P, p.m) == MARCH_HIT ? ST_HIT : ST_MISS;
            } else if (begun || st == ST_SLOW) {
#if defined(RT_JIT_TSTOP)
                if (begun) p.m.t_stop = ray_t_stop<VAR>(P, p.m);
#endif
#if defined(RT_JIT_FAST)
                if (MODE == MODE_HITS) {
                    // hit batches carry no full-code march steps and no region test: a bounce that begins outside the
                    // fast region (rare: bounces start on surfaces inside it) drops out of the march loop at its first
                    // step and then waits for a regeneration batch like every other drop-out
                } else if (MODE == MODE_FRESH && RT_SLOW_FIRST) {
                    st = ST_SLOW;                 // a new primary ray: its steps towards the region at the top of the next iteration
                } else {
                    const int pre = slow_march<VAR>(P, p.m);
                    st = pre == MARCH_CONTINUE ? ST_READY : (pre == MARCH_HIT ? ST_HIT : ST_MISS);
                }
#endif
            }
        }
        const bool more = MODE == MODE_FRESH ? (st == ST_DONE || st == ST_FETCH || st == ST_MISS || (RT_SLOW_FIRST && st == ST_SLOW))   // (hits go to the pending stack)
                        : MODE == MODE_HITS  ? (st == ST_HIT || st == ST_MISS || st == ST_DONE)
                        : st == ST_HIT || st == ST_MISS || (VAR::FAMILY == FAMILY_C ? (st == ST_ADVANCE) : (st == ST_DONE));
#else
        const bool more = MODE == MODE_FRESH ? (st == ST_DONE || st == ST_FETCH)
                        : MODE == MODE_HITS  ? (st == ST_DONE)
                        : VAR::FAMILY == FAMILY_C ? (st == ST_ADVANCE) : (st == ST_DONE);
#endif
        if (__ballot_sync(kFull, more) == 0u) break;
    }

    // ---- write the slot back; ready slots go on the ready stack
    if (slot >= 0) {
        pool.seti(F_STATUS, slot, st);
        const bool marched_here = MODE == MODE_FRESH && (st == ST_HIT || st == ST_MISS);   // marched right here (irregular ray, or full-code steps)
        if (marched_here) pool.setf(F_TEVAL, slot, p.m.t_eval);
        if (st == ST_READY || marched_here || (MODE == MODE_HITS && st == ST_SLOW)) {
            store_ready<VAR, NSLOT>(pool, slot, p.m);
            pool.setf(F_COLX, slot, p.col.x); pool.setf(F_COLY, slot, p.col.y); pool.setf(F_COLZ, slot, p.col.z);
            pool.seti(F_DEPTH, slot, p.depth);
            pool.seti(F_PIXEL, slot, (int)pixel);
            pool.seti(F_SAMP, slot, samp);
            pool.seti(F_K, slot, k);
            pool.seti(F_RNGN, slot, (int)p.rng.n);
            if (VAR::FAMILY == FAMILY_C) {
                pool.setf(F_ACCX, slot, acc.x); pool.setf(F_ACCY, slot, acc.y);
                pool.setf(F_ACCZ, slot, acc.z); pool.setf(F_ACCW, slot, acc.w);
            } else {
                pool.seti(F_ACCX, slot, (int)(uint32_t)wid);
                pool.seti(F_ACCY, slot, (int)(uint32_t)(wid >> 32));
            }
        }
    }
    const unsigned rdy = __ballot_sync(kFull, slot >= 0 && st == ST_READY);
    if (slot >= 0 && st == ST_READY) ready[n_ready + __popc(rdy & lane_lt)] = (uint8_t)slot;
    n_ready += __popc(rdy);
    if (MODE == MODE_HITS) {          // ended paths (and bounces that begin outside the fast region): onto the fresh stack
        const bool to_fresh = slot >= 0 && (st == ST_FETCH || st == ST_SLOW);
        const unsigned fr = __ballot_sync(kFull, to_fresh);
        if (to_fresh) fresh[*n_fresh + __popc(fr & lane_lt)] = (uint8_t)slot;
        *n_fresh += __popc(fr);
    }
    if (MODE == MODE_FRESH) {         // irregular new rays were marched right here: they need a hit / miss batch
        const unsigned hm = __ballot_sync(kFull, slot >= 0 && (st == ST_HIT || st == ST_MISS));
        if (slot >= 0 && (st == ST_HIT || st == ST_MISS)) pend[*n_pend + __popc(hm & lane_lt)] = (uint8_t)slot;
        *n_pend += __popc(hm);
    }
    __syncwarp();
}

template <class VAR, int NSLOT>
__device__ __forceinline__ void pool_body(const KParams& P)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lane_lt = (1u << lane) - 1u;
    constexpr int kWarpWords = F_COUNT * NSLOT + kPoolStacks * (NSLOT / 4) + kPoolQueueWords;
    constexpr bool kRegen = RT_REGEN_MIN > 0 && VAR::FAMILY != FAMILY_C;
    Pool<NSLOT> pool;
    pool.w = smem + warp * kWarpWords;
    uint8_t* ready = reinterpret_cast<uint8_t*>(pool.w + F_COUNT * NSLOT);
    uint8_t* pend = ready + NSLOT;
    uint8_t* fresh = pend + NSLOT;            // kRegen: slots waiting for a regeneration batch (a new path, or full-code march steps)
    volatile uint32_t* wq = pool.w + F_COUNT * NSLOT + kPoolStacks * (NSLOT / 4);   // the warp's chunk of the work queue
    int n_ready = 0, n_pend = kRegen ? 0 : NSLOT, n_fresh = kRegen ? NSLOT : 0;          // warp-uniform
    if (lane < kPoolQueueWords) wq[lane] = 0u;

    for (int s = lane; s < NSLOT; s += 32) {
        pool.seti(F_STATUS, s, ST_FETCH);
        (kRegen ? fresh : pend)[s] = (uint8_t)s;
    }
    __syncwarp();

    int my = -1;
    MarchState m;
    zero_march(m);
    WorkCounters cnt = { 0, 0, 0, 0 };
    unsigned long long c_iters = 0, c_active = 0, c_rounds = 0, c_resolved = 0;

    unsigned active = 0u;   // warp-uniform: lanes that hold a slot (my >= 0)
#if defined(RT_JIT_SPLIT_BUNNY)
    static_assert(VAR::MARCHER == MARCH_ENHANCED && VAR::SHAPESET == SHAPESET_BUNNY, "jit_codegen.h: split march");
    bool need = false;      // this lane stands at a point whose distance still needs the bunny MLP
    float cheap = 0.0f;     // ... the minimum over everything else at that point
    vec3 pb = V3(0.0f);     // ... and the point in the bunny's frame
#endif
    for (;;) {
        // ---------------------------------------------------------------- acquire ready slots
        // (after a resolve round, or when the refill below ran the ready stack dry earlier)
        if (active != kFull && n_ready > 0) {
            const unsigned needy = ~active;
            const int r = __popc(needy & lane_lt);
            if (my < 0 && r < n_ready) {
                my = ready[n_ready - 1 - r];
                load_march<VAR, NSLOT>(pool, my, m);
            }
            n_ready -= min(__popc(needy), n_ready);
            active = __ballot_sync(kFull, my >= 0);
        }

        // ---------------------------------------------------------------- resolve phase
        int n_wait = n_pend;          // slots the resolve phase could work on
        bool regen = false;           // warp-uniform: regeneration batches run in this round
        if (kRegen && active != kFull) {
            regen = n_fresh > 0 && (n_fresh >= RT_REGEN_MIN || 32 - __popc(active) >= RT_REGEN_IDLE);
            if (regen) n_wait += n_fresh;
        }
        if (active != kFull && n_wait > 0 && (n_wait >= P.resolve_min || active == 0u)) {
            if (VAR::COUNT) c_rounds++;
            __syncwarp();    // the acquire above may have read the stack entries that parking overwrites
            if (my >= 0) {   // park: the slot stays ready-to-march
                store_parked<VAR, NSLOT>(pool, my, m);
#if defined(RT_JIT_SPLIT_BUNNY)
                need = false;   // the pending evaluation is simply redone when the slot marches again
#endif
                ready[n_ready + __popc(active & lane_lt)] = (uint8_t)my;
                my = -1;
                idle_march(m);
            }
            n_ready += __popc(active);
            active = 0u;
            __syncwarp();
            // One batch always; further batches only while they are full.  A small remainder stays on
            // the pending stack for the next round instead of costing a whole 32-wide pass.
            if (!kRegen) {
                do {
                    const int take = min(n_pend, 32);
                    const int slot = lane < take ? (int)pend[n_pend - 1 - lane] : -1;
                    n_pend -= take;
                    if (VAR::COUNT) c_resolved += (unsigned long long)take;
                    resolve_batch<VAR, NSLOT>(P, pool, slot, lane, lane_lt, ready, n_ready, wq, cnt);
                } while (n_pend >= 32);
            } else {
                while (n_pend > 0) {
                    const int take = min(n_pend, 32);
                    const int slot = lane < take ? (int)pend[n_pend - 1 - lane] : -1;
                    n_pend -= take;
                    if (VAR::COUNT) c_resolved += (unsigned long long)take;
                    resolve_batch<VAR, NSLOT, MODE_HITS>(P, pool, slot, lane, lane_lt, ready, n_ready, wq, cnt, pend, &n_pend, fresh, &n_fresh);
                    if (n_pend < 32) break;
                }
                // regeneration: when enough slots wait for a new path, or when lanes would idle otherwise
                for (;;) {
                    if (n_fresh == 0 || !(regen || n_fresh >= RT_REGEN_MIN)) break;
                    const int take = min(n_fresh, 32);
                    const int slot = lane < take ? (int)fresh[n_fresh - 1 - lane] : -1;
                    n_fresh -= take;
                    __syncwarp();
                    resolve_batch<VAR, NSLOT, MODE_FRESH>(P, pool, slot, lane, lane_lt, ready, n_ready, wq, cnt, pend, &n_pend, fresh, &n_fresh);
                    regen = false;                                // further batches only while they are well filled
                }
            }
            continue;
        }
        if (active == 0u) break;   // nothing marching, nothing pending, nothing ready: pool drained

        // ---------------------------------------------------------------- march loop
        // One sphere-tracing step per iteration for every lane; lanes without a slot march a harmless
        // dummy ray (idle_march) and are masked out of the vote, which keeps the loop body branch-free.
        unsigned fin;
        float aux;
        bool slow = false;   // fast-region kernels: the lane's next evaluation point lies outside the region
#if defined(RT_JIT_SPLIT_BUNNY)
        // Scenes with the neural bunny: a step is cheap outside the bunny's unit sphere (|p| - 0.8 and the
        // analytic objects) and ~1700 instructions inside (the MLP).  Marching them in lockstep left a third
        // of the lanes idle during the MLP (ncu: 22.4 of 32 threads active there), so the loop has two
        // stages: each lane first takes all the cheap steps it can -- until it ends or stands at a point that
        // needs the MLP -- and then the MLP runs for every marching lane at once.  Per ray the sequence of
        // evaluations is unchanged; only the interleaving across lanes is.
        for (;;) {
            bool f = false;
            while (my >= 0 && !need && !f) {
                cheap = jit_nearest_partial(P, at(m.ro, m.rd, m.t), need, pb);
                if (!need) f = enhanced_advance(P, m, cheap, aux);
            }
            fin = __ballot_sync(kFull, f);
            if (fin != 0u) break;
            if (need) {
                const float dist = fminf(cheap, fabsf(sd_bunny_mlp(pb.x, pb.y, pb.z)));
                need = false;
                f = enhanced_advance(P, m, dist, aux);
            }
            fin = __ballot_sync(kFull, f);
            if (fin != 0u) break;
        }
#else
#if RT_FIN_MIN > 1
        // Leaving the march loop costs ~45 warp instructions (store, push, pop, load) however few lanes finished, and
        // with 32 lanes x ~28 steps per ray some lane finishes on almost every step.  So finished lanes wait (masked)
        // until RT_FIN_MIN of them can be handled in one go -- or every lane that holds a slot has finished.
        {
            const int thr = min(RT_FIN_MIN, __popc(__ballot_sync(kFull, my >= 0)));
            bool done = false;
            do {
                if (!done) done = march_step_fin<VAR>(P, m, aux, slow);
                fin = __ballot_sync(kFull, done && my >= 0);
            } while (__popc(fin) < thr);
        }
#elif defined(RT_MARCH_VOTE_EVERY_2)
        // tuning knob (RTPBR_MARCH_UNROLL=2): vote after every second step; a lane whose march ends on the first
        // one sits the second one out
        do {
            bool f = march_step_fin<VAR>(P, m, aux, slow);
            if (!f) f = march_step_fin<VAR>(P, m, aux, slow);
            fin = __ballot_sync(kFull, f && my >= 0);
        } while (fin == 0u);
#else
        do {
            if (VAR::COUNT) { c_iters += 32; c_active += (unsigned long long)__popc(active); }
            const bool f = march_step_fin<VAR>(P, m, aux, slow);
            fin = __ballot_sync(kFull, f && my >= 0);
        } while (fin == 0u);
#endif
#endif

        // ---------------------------------------------------------------- finished lanes: push the slot on the
        // pending stack and take a ready-to-march one straight away (no extra pass while the ready stack lasts)
        {
            const int nf = __popc(fin);
            const bool mine = (fin >> lane) & 1u;
#if defined(RT_JIT_FAST)
            const bool dropped = mine && march_undo_slow<VAR>(m, aux, slow);   // nothing was advanced: the march goes on in the resolve phase
#else
            const bool dropped = false;
#endif
            const bool hit = mine && !dropped && march_status<VAR>(P, aux) == MARCH_HIT;
            // with regeneration batches the drop-outs wait for one of those (full-code steps at full lanes) instead of
            // riding along in a hit batch, and so do the missed rays: their path ends, all that is left to do is the
            // regeneration, and the hit batches stay free of lanes that sit out the shading
            const unsigned drop_mask = kRegen ? __ballot_sync(kFull, mine && !hit) : 0u;
            if (mine) {
                const int r = __popc(fin & lane_lt);
                if (dropped) {
                    store_parked<VAR, NSLOT>(pool, my, m);
                    pool.seti(F_STATUS, my, ST_SLOW);
                } else {
                    store_finished<VAR, NSLOT>(pool, my, m, hit ? ST_HIT : ST_MISS);
                }
                if (kRegen && !hit) fresh[n_fresh + __popc(drop_mask & lane_lt)] = (uint8_t)my;
                else pend[n_pend + __popc(fin & ~drop_mask & lane_lt)] = (uint8_t)my;
                if (r < n_ready) {
                    my = ready[n_ready - 1 - r];
                    load_march<VAR, NSLOT>(pool, my, m);
                } else {
                    my = -1;
                    idle_march(m);
                }
            }
            n_pend += nf - __popc(drop_mask);
            n_fresh += __popc(drop_mask);
            if (nf <= n_ready) {
                n_ready -= nf;
            } else {
                n_ready = 0;
                active = __ballot_sync(kFull, my >= 0);
            }
            __syncwarp();
        }
    }

    if (VAR::COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cnt.evals += __shfl_xor_sync(kFull, cnt.evals, o);
            cnt.rays += __shfl_xor_sync(kFull, cnt.rays, o);
            cnt.normals += __shfl_xor_sync(kFull, cnt.normals, o);
            cnt.samples += __shfl_xor_sync(kFull, cnt.samples, o);
        }
        if (lane == 0) {
            atomicAdd(&P.counters[0], cnt.evals);
            atomicAdd(&P.counters[1], cnt.rays);
            atomicAdd(&P.counters[2], cnt.normals);
            atomicAdd(&P.counters[3], cnt.samples);
            atomicAdd(&P.counters[4], c_iters);
            atomicAdd(&P.counters[5], c_active);
            atomicAdd(&P.counters[6], c_rounds);
            atomicAdd(&P.counters[8], c_resolved);
        }
    }
}


// The neural bunny's out-of-line MLP keeps ~50 floats live: its kernels get more registers per thread
// (fewer resident CTAs) than the analytic scenes.  RT_POOL_MIN_BLOCKS_BUNNY overrides (NVRTC builds: env
// RTPBR_POOL_MIN_BLOCKS_BUNNY, capi.cu).
#ifndef RT_POOL_MIN_BLOCKS_BUNNY
#define RT_POOL_MIN_BLOCKS_BUNNY 2
#endif
template <class VAR>
struct PoolLaunch {
    static constexpr int kMinBlocks = VAR::SHAPESET == SHAPESET_BUNNY ? RT_POOL_MIN_BLOCKS_BUNNY : kPoolMinBlocks;
};

template <class VAR, int NSLOT>
__global__ void __launch_bounds__(kPoolBlock, PoolLaunch<VAR>::kMinBlocks) k_pathtrace_pool(const __grid_constant__ KParams P)
{
    pool_body<VAR, NSLOT>(P);
}

}  // namespace rt
