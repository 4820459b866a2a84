// C-ABI of the Fetch scenario (include/mplb.h: mplb_scene_create_fetch*): replaces FetchScenario's
// constructor and validity methods (reference include/mpl/demo/fetch_scenario.hpp:48-63,105-173).
#include "../../include/mplb.h"
#include "fetch_device.cuh"
#include "mplb_host.hpp"
#include "scene_impl.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

namespace mplb {
namespace {

// ---- Origin(xyz, rpy) of fetch_robot.hpp:23-35 with Eigen's arithmetic ([EXT]): AngleAxis products go
// through quaternions, then Quaternion::toRotationMatrix.  Same operation order as the CPU path.
void quatFromAA(double angle, int axis, double q[4]) {
    const double h = 0.5 * angle, s = std::sin(h);
    q[0] = q[1] = q[2] = 0;
    q[axis] = s;
    q[3] = std::cos(h);
}
void quatMul(const double a[4], const double b[4], double r[4]) {
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
void originFrame(double tx, double ty, double tz, double roll, double pitch, double yaw, double out[12]) {
    double qz[4], qy[4], qx[4], a[4], q[4];
    quatFromAA(yaw, 2, qz);
    quatFromAA(pitch, 1, qy);
    quatFromAA(roll, 0, qx);
    quatMul(qz, qy, a);
    quatMul(a, qx, q);
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx2 = 2 * x, ty2 = 2 * y, tz2 = 2 * z;
    const double twx = tx2 * w, twy = ty2 * w, twz = tz2 * w;
    const double txx = tx2 * x, txy = ty2 * x, txz = tz2 * x;
    const double tyy = ty2 * y, tyz = tz2 * y, tzz = tz2 * z;
    out[0] = 1 - (tyy + tzz); out[1] = txy - twz;       out[2] = txz + twy;        out[3] = tx;
    out[4] = txy + twz;       out[5] = 1 - (txx + tzz); out[6] = tyz - twx;        out[7] = ty;
    out[8] = txz - twy;       out[9] = tyz + twx;       out[10] = 1 - (txx + tyy); out[11] = tz;
}

// edge items per round (frames scratch: 144 doubles per item); MPLB_FETCH_EDGE_BATCH overrides (diagnostic)
static const size_t kBatch = [] {
    const char *e = std::getenv("MPLB_FETCH_EDGE_BATCH");
    const long v = e ? std::atol(e) : 0;
    return v >= 1024 ? (size_t)v : (size_t)1 << 16;
}();
// states per launch group of a state batch (151 MB of frames); MPLB_FETCH_BATCH overrides (diagnostic).  Measured for 2^18
// configurations end to end: 32 K 2.52 ms, 64 K 1.81 ms, 128 K 1.56 ms, 256 K 1.58 ms -- six kernels per group, each with
// a tail; beyond 128 K the first group's input copy is no longer hidden.  (Edge rounds stay at 64 K items: 128 K measured
// 9.1 -> 16.8 ms for 16,384 edges.)
static const size_t kStateBatch = [] {
    const char *e = std::getenv("MPLB_FETCH_BATCH");
    const long v = e ? std::atol(e) : 0;
    return v >= 1024 ? (size_t)v : (size_t)1 << 17;
}();

}  // namespace

struct FetchScene : SceneBase {
    double invStep = 100;
    FetchSceneDev dev{};
    DevBuf envNodes, envTris64;
    size_t nEnvTris = 0;
    std::atomic<uint64_t> hPairTests{0}, hLeafTests{0};

    FetchScene() { stateDim = 8; }

    // nigh L1Space<S,8>, unscaled (fetch_scenario.hpp:127)
    double distance(const double *a, const double *b) const override {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += std::fabs(a[i] - b[i]);
        return s;
    }
    uint64_t edgeSteps(const double *a, const double *b) const {
        const double x = std::ceil(distance(a, b) * invStep);   // (non-finite: rejected by the callers, never converted)
        return x < 9.0e18 ? (uint64_t)x : ~0ull;
    }
    int planEdges(const double *from, const double *to, size_t n, uint32_t *steps) const override {
        for (size_t i = 0; i < n; ++i) {
            const uint64_t st = edgeSteps(from + 8 * i, to + 8 * i);
            if (!(st < (1ull << 31))) return fail(MPLB_ERR_INVALID, "edge %zu: step count out of range", i);
            steps[i] = (uint32_t)st;
        }
        return MPLB_OK;
    }

    // Work lists between the culling and the exact kernels: sized for the average (8 pairs, 16 leaf pairs per state; a
    // state lists ~1.3).  A full pair list degrades gracefully (the pair is decided on the spot); a full leaf list drops
    // leaf pairs and flags the group (FetchCounters::pad) -- the call is then repeated with `safe` lists that hold the
    // worst case, every geometry against every triangle, in groups of safeGroup() states.
    size_t envTris() const { return (envNodeCount + 1) / 2; }
    size_t safeGroup() const {
        const size_t perState = 12 * std::max<size_t>(envTris(), 1) * 8;
        return std::max<size_t>(16, std::min<size_t>(8192, ((size_t)64 << 20) / perState));
    }
    int reserveScratch(CallCtx *c, size_t n, FetchScratch &w, bool safe = false) {
        int rc;
        const size_t pc = safe ? 24 * n : fetchPairCap(n), lc = safe ? 12 * n * std::max<size_t>(envTris(), 1) : fetchLeafCap(n);
        // aux4: the pair list, and behind it the pairs that need MPR; aux5: the leaf list
        if ((rc = c->aux0.reserve(n * 144 * sizeof(double))) || (rc = c->aux1.reserve(n)) ||
            (rc = c->aux2.reserve(n * sizeof(unsigned))) || (rc = c->aux4.reserve(2 * pc * 8)) ||
            (rc = c->aux5.reserve(lc * 8)) || (rc = c->aux6.reserve(n * sizeof(unsigned))) ||
            (rc = c->work.reserve(64)) || (rc = c->out.reserve(n)))
            return rc;
        w.frames = (double *)c->aux0.ptr;
        w.selfHit = (unsigned char *)c->aux1.ptr;
        w.envHit = (unsigned *)c->aux2.ptr;
        w.survivors = (unsigned *)c->aux6.ptr;
        w.pairItems = (unsigned long long *)c->aux4.ptr;
        w.leafItems = (unsigned long long *)c->aux5.ptr;
        // MPLB_FETCH_LIST_CAP (diagnostic, tests): tiny work lists, so that the overflow paths run
        static const size_t listCap = [] {
            const char *e = std::getenv("MPLB_FETCH_LIST_CAP");
            return e ? (size_t)std::max(1L, std::atol(e)) : ~(size_t)0;
        }();
        w.pairCap = safe ? pc : std::min(pc, listCap);
        w.leafCap = safe ? lc : std::min(lc, listCap);
        w.counts = (unsigned *)c->work.ptr;
        w.touched = (unsigned char *)c->out.ptr;   // (checkStates reuses the buffer for its verdicts afterwards)
        return MPLB_OK;
    }

    static constexpr int kListOverflow = 1;   // finish(): a leaf list overflowed, the call has to be repeated with safe lists
    int finish(CallCtx *c) {
        FetchCounters h{};
        cudaError_t e = cudaMemcpyAsync(&h, c->ctr.ptr, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(MPLB_ERR_CUDA, "device work failed: %s", cudaGetErrorString(e));
        if (h.error) return fail(MPLB_ERR_CUDA, "device traversal bookkeeping error (Fetch stack / MPR iteration guard)");
        hBv += h.bvTests;
        hPairTests += h.pairTests;
        hLeafTests += h.leafTests;
        hTri += h.leafTests;
        return h.pad ? kListOverflow : MPLB_OK;
    }

    int checkStates(const double *states, size_t n, uint8_t *valid) override {
        int rc = checkStatesImpl(states, n, valid, false);
        if (rc == kListOverflow) {
            hListOverflows += 1;
            rc = checkStatesImpl(states, n, valid, true);
            if (rc == kListOverflow) return fail(MPLB_ERR_CUDA, "Fetch leaf list overflowed although it holds the worst case");
        }
        return rc;
    }
    std::atomic<uint64_t> hListOverflows{0};

    int checkStatesImpl(const double *states, size_t n, uint8_t *valid, bool safe) {
        MPLB_CUDA(cudaSetDevice(device));
        CallCtx *c = acquire();
        if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
        CtxGuard g{this, c};
        const size_t kStateBatch = safe ? safeGroup() : ::mplb::kStateBatch;
        const size_t nb = std::min(n, kStateBatch);
        int rc;
        FetchScratch w{};
        // chunks of kStateBatch states: the input of chunk i + 1 crosses the link (copy stream, second buffer) while chunk i
        // is being checked; the verdicts of all chunks go back in one copy at the end
        if ((rc = c->in0.reserve(nb * 8 * sizeof(double))) || (rc = c->in1.reserve(nb * 8 * sizeof(double))) ||
            (rc = c->dead.reserve(n)) || (rc = reserveScratch(c, nb, w, safe)))
            return rc;
        while (c->events.size() < 4) {
            cudaEvent_t ev;
            MPLB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->events.push_back(ev);
        }
        MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(FetchCounters), c->stream));
        size_t chunk = 0;
        for (size_t lo = 0; lo < n; lo += kStateBatch, ++chunk) {
            const size_t cnt = std::min(kStateBatch, n - lo);
            const int b = (int)(chunk & 1);
            double *in = (double *)(b ? c->in1.ptr : c->in0.ptr);
            if (chunk >= 2) MPLB_CUDA(cudaStreamWaitEvent(c->copyStream, c->events[2 + b], 0));   // the buffer's previous chunk is done
            if ((rc = hostToDevice(c, in, states + 8 * lo, cnt * 8 * sizeof(double), c->copyStream))) return rc;
            MPLB_CUDA(cudaEventRecord(c->events[b], c->copyStream));
            MPLB_CUDA(cudaStreamWaitEvent(c->stream, c->events[b], 0));
            MPLB_CUDA(launchFetchStates(dev, in, cnt, w, (uint8_t *)c->dead.ptr + lo, (FetchCounters *)c->ctr.ptr, numSms,
                                        c->stream));
            MPLB_CUDA(cudaEventRecord(c->events[2 + b], c->stream));
            hLaunches += 6;
        }
        MPLB_CUDA(cudaMemcpyAsync(valid, c->dead.ptr, n, cudaMemcpyDeviceToHost, c->stream));
        hChecks += n;
        rc = finish(c);
        cudaStreamSynchronize(c->copyStream);
        return rc;
    }

    // isValid(from,to) (fetch_scenario.hpp:120-173): verdict = valid(to) AND every interpolated sample.  The samples
    // are not enumerated: an item is `to`, or an INTERVAL [lo, hi] of samples evaluated at its middle one with every
    // box test inflated by how far each geometry can stray over the interval (FetchSceneDev::reach).  An interval
    // that nothing comes within reach of is certified free as a whole; otherwise its middle sample has been tested
    // exactly and the two halves go into the next round.  Every sample is the middle of exactly one interval, so the
    // verdict is the reference's discrete one.  Rounds are host driven (the work lists live on the host).
    static constexpr int kIntervalOverflow = 2;   // checkEdgesDevice: the device-side interval lists did not hold a round
    int checkEdges(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked) override {
        static const bool hostRounds = std::getenv("MPLB_FETCH_HOST_ROUNDS") != nullptr;   // diagnostic: round 1's host-side lists
        bool onHost = hostRounds || n >= ((size_t)1 << 31);
        int rc = kIntervalOverflow;
        for (int safe = 0; safe < 2; ++safe) {
            if (!onHost) rc = checkEdgesDevice(from, to, n, valid, statesChecked, safe != 0);
            if (onHost || rc == kIntervalOverflow) {
                onHost = true;
                rc = checkEdgesImpl(from, to, n, valid, statesChecked, safe != 0);
            }
            if (rc != kListOverflow) return rc;
            hListOverflows += 1;
        }
        return fail(MPLB_ERR_CUDA, "Fetch leaf list overflowed although it holds the worst case");
    }

    // isValid(from,to) with the interval lists on the device (fetch_kernels.cu: fetchRound0Kernel .. fetchSplitKernel): per
    // round the host launches the groups -- items from intervals, the six validity kernels, the split of what was neither
    // certified nor killed into the next round's list -- and reads the next round's length back.  Same intervals, same
    // tests and the same verdicts as the host-side rounds below, which remain as the fallback for a round that outgrows
    // the device lists.
    int checkEdgesDevice(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked, bool safe) {
        const size_t B = safe ? safeGroup() : ::mplb::kBatch;   // items per launch group
        MPLB_CUDA(cudaSetDevice(device));
        CallCtx *c = acquire();
        if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
        CtxGuard g{this, c};
        c->hSteps.resize(n);
        int rc = planEdges(from, to, n, c->hSteps.data());
        if (rc != MPLB_OK) return rc;
        const size_t sb = n * 8 * sizeof(double);
        const size_t listCap = std::min<size_t>(std::max<size_t>(17 * n + 1024, 64 * n), (size_t)1 << 26);
        const size_t offForced = (n * sizeof(float2) + 255) & ~(size_t)255, offCounts = (offForced + B + 255) & ~(size_t)255;
        FetchScratch w{};
        if ((rc = c->in0.reserve(sb)) || (rc = c->in1.reserve(sb)) || (rc = c->steps.reserve(n * 4)) ||
            (rc = c->dead.reserve(n * 4)) || (rc = c->aux3.reserve(B * 8)) || (rc = c->stage.reserve(B * sizeof(float2))) ||
            (rc = c->ivA.reserve(listCap * sizeof(uint4))) || (rc = c->ivB.reserve(listCap * sizeof(uint4))) ||
            (rc = c->ivMisc.reserve(offCounts + 64)) || (rc = reserveScratch(c, B, w, safe)))
            return rc;
        uint4 *list[2] = {(uint4 *)c->ivA.ptr, (uint4 *)c->ivB.ptr};
        float2 *motion = (float2 *)c->ivMisc.ptr;
        unsigned char *forced = (unsigned char *)c->ivMisc.ptr + offForced;
        unsigned *counts = (unsigned *)((unsigned char *)c->ivMisc.ptr + offCounts);   // [0], [1]: list lengths, [2]: overflow
        volatile unsigned *hCounts = (volatile unsigned *)c->hWater;                 // pinned
        static const float kCap = std::getenv("MPLB_FETCH_CAP") ? (float)std::atof(std::getenv("MPLB_FETCH_CAP")) : 0.2f;
        MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(FetchCounters), c->stream));
        MPLB_CUDA(cudaMemsetAsync(c->dead.ptr, 0, n * 4, c->stream));
        MPLB_CUDA(cudaMemsetAsync(counts, 0, 16, c->stream));
        int hrc;
        if ((hrc = hostToDevice(c, c->in0.ptr, from, sb, c->stream)) || (hrc = hostToDevice(c, c->in1.ptr, to, sb, c->stream))) return hrc;
        MPLB_CUDA(cudaMemcpyAsync(c->steps.ptr, c->hSteps.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        MPLB_CUDA(launchFetchEdgeMotion((const double *)c->in0.ptr, (const double *)c->in1.ptr, n, motion, c->stream));
        MPLB_CUDA(launchFetchRound0((const unsigned *)c->steps.ptr, n, list[0], counts, (unsigned)listCap, counts + 2, c->stream));
        uint64_t launched = 0;
        for (int cur = 0;; cur ^= 1) {
            MPLB_CUDA(cudaMemcpyAsync((void *)hCounts, counts, 16, cudaMemcpyDeviceToHost, c->stream));
            MPLB_CUDA(cudaStreamSynchronize(c->stream));
            if (hCounts[2]) return kIntervalOverflow;
            const size_t cnt = hCounts[cur];
            if (cnt == 0) break;
            MPLB_CUDA(cudaMemsetAsync(counts + (cur ^ 1), 0, sizeof(unsigned), c->stream));
            // small rounds are all latency: cut what is left into more pieces per interval, for fewer rounds
            static const unsigned smallRound = [] {
                const char *e = std::getenv("MPLB_FETCH_SMALL_ROUND");   // diagnostic
                return e ? (unsigned)std::atol(e) : 65536u;
            }();
            const unsigned minParts = cnt < smallRound / 4 ? 4u : cnt < smallRound ? 2u : 1u;
            for (size_t first = 0; first < cnt; first += B) {
                const size_t m = std::min(B, cnt - first);
                MPLB_CUDA(launchFetchItems(list[cur] + first, m, (const unsigned *)c->steps.ptr, motion, kCap,
                                           (unsigned long long *)c->aux3.ptr, (float2 *)c->stage.ptr, forced, c->stream));
                MPLB_CUDA(launchFetchEdgeItems(dev, (const double *)c->in0.ptr, (const double *)c->in1.ptr,
                                               (const unsigned *)c->steps.ptr, (const unsigned long long *)c->aux3.ptr,
                                               (const float2 *)c->stage.ptr, m, (unsigned *)c->dead.ptr, w,
                                               (FetchCounters *)c->ctr.ptr, numSms, c->stream));
                MPLB_CUDA(launchFetchSplit(list[cur] + first, m, w.touched, forced, (const unsigned *)c->dead.ptr,
                                           (const unsigned *)c->steps.ptr, motion, kCap, minParts, list[cur ^ 1], counts + (cur ^ 1),
                                           (unsigned)listCap, counts + 2, c->stream));
                hLaunches += 8;
                launched += m;
            }
        }
        c->hSteps.resize(n);   // (reused as the landing buffer of the dead flags)
        MPLB_CUDA(cudaMemcpyAsync(c->hSteps.data(), c->dead.ptr, n * 4, cudaMemcpyDeviceToHost, c->stream));
        rc = finish(c);
        if (rc != MPLB_OK) return rc;
        for (size_t e = 0; e < n; ++e) valid[e] = c->hSteps[e] ? 0 : 1;
        hChecks += launched;
        if (statesChecked) *statesChecked = launched;
        return MPLB_OK;
    }
    int checkEdgesImpl(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked, bool safe) {
        const size_t kBatch = safe ? safeGroup() : ::mplb::kBatch;   // items per round
        MPLB_CUDA(cudaSetDevice(device));
        CallCtx *c = acquire();
        if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
        CtxGuard g{this, c};
        c->hSteps.resize(n);
        int rc = planEdges(from, to, n, c->hSteps.data());
        if (rc != MPLB_OK) return rc;
        const size_t sb = n * 8 * sizeof(double);
        FetchScratch w{};
        if ((rc = c->in0.reserve(sb)) || (rc = c->in1.reserve(sb)) || (rc = c->steps.reserve(n * 4)) ||
            (rc = c->dead.reserve(n * 4)) || (rc = c->aux3.reserve(kBatch * 8)) || (rc = c->stage.reserve(kBatch * sizeof(float2))) ||
            (rc = reserveScratch(c, kBatch, w, safe)))
            return rc;
        MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(FetchCounters), c->stream));
        MPLB_CUDA(cudaMemsetAsync(c->dead.ptr, 0, n * 4, c->stream));
        MPLB_CUDA(cudaMemcpyAsync(c->in0.ptr, from, sb, cudaMemcpyHostToDevice, c->stream));
        MPLB_CUDA(cudaMemcpyAsync(c->in1.ptr, to, sb, cudaMemcpyHostToDevice, c->stream));
        MPLB_CUDA(cudaMemcpyAsync(c->steps.ptr, c->hSteps.data(), n * 4, cudaMemcpyHostToDevice, c->stream));

        struct Interval { uint32_t e, lo, hi; };   // lo == hi == steps: the edge's `to`
        std::vector<Interval> cur, next;
        std::vector<unsigned long long> items;
        std::vector<float2> slack;
        std::vector<unsigned> dead(n, 0);
        uint64_t launched = 0;
        // joint motion over the whole edge: torso (prismatic) and the arm joints (L1)
        std::vector<float> dTorso(n), dArm(n);
        for (size_t e = 0; e < n; ++e) {
            double a = 0;
            for (int j = 1; j < 8; ++j) a += std::fabs(to[8 * e + j] - from[8 * e + j]);
            dTorso[e] = (float)(std::fabs(to[8 * e] - from[8 * e]) * 1.000001);
            dArm[e] = (float)(a * 1.000001);
        }
        // Two batches are in flight: while the device works on one, the host turns the results of the previous one
        // into next-round intervals and builds the following one.  Results land in pinned buffers (slot = batch & 1).
        const size_t resBytes = kBatch + n * 4;
        for (int i = 0; i < 2; ++i)
            if ((rc = c->reserveRes(i, resBytes))) return rc;
        while (c->events.size() < 2) {
            cudaEvent_t ev;
            MPLB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->events.push_back(ev);
        }
        static const float kCap = std::getenv("MPLB_FETCH_CAP") ? (float)std::atof(std::getenv("MPLB_FETCH_CAP")) : 0.2f;
        std::vector<Interval> inflight[2];
        std::vector<unsigned char> forced[2] = {std::vector<unsigned char>(kBatch), std::vector<unsigned char>(kBatch)};
        unsigned long long dbgHist[16][32] = {};
        auto submit = [&](int slot) -> int {
            std::vector<Interval> &b = inflight[slot];
            items.resize(b.size());
            slack.resize(b.size());
            // (the host side of a round is as much work as the device side: all cores)
#pragma omp parallel for schedule(static) if (b.size() >= 4096)
            for (long long i = 0; i < (long long)b.size(); ++i) {
                const Interval &v = b[i];
                const uint32_t st = c->hSteps[v.e], mid = v.lo + (v.hi - v.lo) / 2;
                items[i] = ((unsigned long long)v.e << 32) | mid;
                float f = 0.f;
                if (v.hi > v.lo) f = (float)((double)std::max(v.hi - mid, mid - v.lo) / (double)st * 1.000001);
                slack[i] = make_float2(dTorso[v.e] * f, dArm[v.e] * f);
                // an interval over which the gripper can stray by more than ~12 cm is practically never certified
                // (measured: 0-2 % from 32 samples up) and its inflated traversal costs more than a plain test:
                // evaluate its middle sample plainly and split it unconditionally
                forced[slot][i] = v.hi > v.lo && slack[i].x + slack[i].y * 1.2f > kCap;
                if (forced[slot][i]) slack[i] = make_float2(0.f, 0.f);
            }
            MPLB_CUDA(cudaMemcpyAsync(c->aux3.ptr, items.data(), items.size() * 8, cudaMemcpyHostToDevice, c->stream));
            MPLB_CUDA(cudaMemcpyAsync(c->stage.ptr, slack.data(), slack.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
            MPLB_CUDA(launchFetchEdgeItems(dev, (const double *)c->in0.ptr, (const double *)c->in1.ptr,
                                           (const unsigned *)c->steps.ptr, (const unsigned long long *)c->aux3.ptr,
                                           (const float2 *)c->stage.ptr, items.size(), (unsigned *)c->dead.ptr, w,
                                           (FetchCounters *)c->ctr.ptr, numSms, c->stream));
            hLaunches += 6;
            launched += items.size();
            MPLB_CUDA(cudaMemcpyAsync(c->hRes[slot], w.touched, items.size(), cudaMemcpyDeviceToHost, c->stream));
            MPLB_CUDA(cudaMemcpyAsync(c->hRes[slot] + kBatch, c->dead.ptr, n * 4, cudaMemcpyDeviceToHost, c->stream));
            MPLB_CUDA(cudaEventRecord(c->events[slot], c->stream));
            return MPLB_OK;
        };
        auto collect = [&](int slot) -> int {
            std::vector<Interval> &b = inflight[slot];
            if (b.empty()) return MPLB_OK;
            MPLB_CUDA(cudaEventSynchronize(c->events[slot]));
            const unsigned char *t = c->hRes[slot];
            std::memcpy(dead.data(), c->hRes[slot] + kBatch, n * 4);
            static const bool dbg = std::getenv("MPLB_FETCH_DEBUG") != nullptr;
#pragma omp parallel if (b.size() >= 4096 && !dbg)
            {
                std::vector<Interval> mine;
#pragma omp for schedule(static) nowait
                for (long long i = 0; i < (long long)b.size(); ++i) {
                    const Interval &v = b[i];
                    if (dbg && !dead[v.e] && v.hi > v.lo) {
                        int cls = 0;
                        for (uint32_t len = v.hi - v.lo + 1; len > 1; len >>= 1) ++cls;
                        ++dbgHist[std::min(cls, 15)][t[i] & 31];
                    }
                    if (dead[v.e] || v.hi <= v.lo || !(t[i] || forced[slot][i])) continue;   // dead, a single sample, or certified
                    const uint32_t mid = v.lo + (v.hi - v.lo) / 2;
                    // a split that was forced goes straight to pieces short enough to stand a chance (<= 2 H + 1 samples,
                    // H the half width at which the gripper's stray reaches the cap), not through log2 plain halvings
                    uint32_t piece = 0xffffffffu;
                    if (forced[slot][i]) {
                        const double perSample = ((double)dArm[v.e] * 1.2 + (double)dTorso[v.e]) / (double)c->hSteps[v.e];
                        const double H = perSample > 0 ? std::floor((double)kCap / perSample) : 1e9;
                        piece = (uint32_t)std::min(1e9, 2.0 * std::max(H, 1.0) + 1.0);
                    }
                    auto cut = [&](uint32_t lo, uint32_t hi) {   // [lo, hi] into equal pieces of at most `piece` samples
                        const uint64_t len = (uint64_t)hi - lo + 1, parts = (len + piece - 1) / piece;
                        for (uint64_t p = 0; p < parts; ++p)
                            mine.push_back({v.e, lo + (uint32_t)(p * len / parts), lo + (uint32_t)((p + 1) * len / parts) - 1});
                    };
                    if (mid > v.lo) cut(v.lo, mid - 1);
                    if (mid < v.hi) cut(mid + 1, v.hi);
                }
#pragma omp critical(mplb_fetch_next)
                next.insert(next.end(), mine.begin(), mine.end());
            }
            b.clear();
            return MPLB_OK;
        };
        // round 0: `to` and up to 16 equal intervals of the samples [1, steps - 1] of every edge
        for (size_t e = 0; e < n; ++e) {
            const uint32_t st = c->hSteps[e];
            cur.push_back({(uint32_t)e, st, st});
            const uint64_t m = st ? st - 1 : 0;
            for (uint64_t k = 0; k < 16; ++k) {
                const uint32_t lo = 1u + (uint32_t)((k * m) >> 4), hi = (uint32_t)(((k + 1) * m) >> 4);
                if (lo <= hi) cur.push_back({(uint32_t)e, lo, hi});
            }
        }
        int slot = 0;
        while (!cur.empty()) {
            size_t pos = 0;
            while (pos < cur.size()) {
                std::vector<Interval> &b = inflight[slot];
                while (pos < cur.size() && b.size() < kBatch) {
                    if (!dead[cur[pos].e]) b.push_back(cur[pos]);
                    ++pos;
                }
                if (b.empty()) break;
                if ((rc = submit(slot)) || (rc = collect(slot ^ 1))) return rc;
                slot ^= 1;
            }
            if ((rc = collect(slot ^ 1)) || (rc = collect(slot))) return rc;   // the round is complete
            cur.swap(next);
            next.clear();
        }
        if (std::getenv("MPLB_FETCH_DEBUG"))
            for (int cls = 0; cls < 16; ++cls) {
                unsigned long long tot = 0, by[5] = {0, 0, 0, 0, 0};
                for (int f = 0; f < 32; ++f) {
                    tot += dbgHist[cls][f];
                    for (int bit = 0; bit < 5; ++bit)
                        if (f >> bit & 1) by[bit] += dbgHist[cls][f];
                }
                if (tot)
                    std::fprintf(stderr, "[fetch intervals] length ~2^%d: %llu, certified %llu; touched by floor %llu, pair near %llu, pair overlap %llu, env near %llu, env overlap %llu\n",
                                 cls, tot, dbgHist[cls][0], by[0], by[1], by[2], by[3], by[4]);
            }
        for (size_t e = 0; e < n; ++e) valid[e] = dead[e] ? 0 : 1;
        hChecks += launched;
        if (statesChecked) *statesChecked = launched;
        return finish(c);
    }
};

static int createFetch(const double *envTris, size_t nEnv, const double envFrame[12], double res, int device,
                       mplb_scene **out) {
    if (!out) return fail(MPLB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (nEnv && !envTris) return fail(MPLB_ERR_INVALID, "null triangle array");
    if (!envFrame) return fail(MPLB_ERR_INVALID, "null environment frame");
    if (!(res > 0)) return fail(MPLB_ERR_INVALID, "check_resolution must be > 0");
    for (size_t i = 0; i < 9 * nEnv; ++i)
        if (!std::isfinite(envTris[i])) return fail(MPLB_ERR_INVALID, "non-finite environment vertex");
    for (int i = 0; i < 12; ++i)
        if (!std::isfinite(envFrame[i])) return fail(MPLB_ERR_INVALID, "non-finite environment frame");
    int numSms = 0;
    int rc = selectDevice(device, &numSms);
    if (rc != MPLB_OK) return rc;
    FlatBvh env = buildBvhCached(envTris, nEnv);
    if (env.depth > 36) return fail(MPLB_ERR_INVALID, "environment hierarchy too deep (%d levels)", env.depth);
    auto sc = std::make_unique<FetchScene>();
    sc->device = device;
    sc->numSms = numSms;
    sc->invStep = 1 / res;
    sc->nEnvTris = nEnv;
    if ((rc = upload(sc->envNodes, env.nodes)) || (rc = upload(sc->envTris64, env.tris64))) return rc;
    if ((rc = sc->globalCtr.reserve(sizeof(DevCounters)))) return rc;
    MPLB_CUDA(cudaMemset(sc->globalCtr.ptr, 0, sizeof(DevCounters)));
    FetchSceneDev &d = sc->dev;
    d.envNodes = (const float4 *)sc->envNodes.ptr;
    d.envTris64 = (const double *)sc->envTris64.ptr;
    d.nEnvNodes = (int)env.nodes.size();
    std::memcpy(d.envFrame, envFrame, sizeof(d.envFrame));
    // fetch_robot.hpp:217-274 (link -> geometry origins), geometry order of fetch_robot.hpp:38-53
    originFrame(0, 0, 0.1975, 0, 0, 0, d.origins + 12 * 0);
    originFrame(-0.09, 0, 0.31, 0, 0, 0, d.origins + 12 * 1);
    originFrame(0, 0, -0.025, 0, 0, 0, d.origins + 12 * 2);
    originFrame(0, -0.0025, 0, 1.5708, 0, 0, d.origins + 12 * 3);
    originFrame(-0.04, 0, 0, 0, 1.5708, 0, d.origins + 12 * 4);
    originFrame(0, 0, 0, 1.5708, 0, 0, d.origins + 12 * 5);
    originFrame(-0.04, 0, 0, 0, 1.5708, 0, d.origins + 12 * 6);
    originFrame(0, -0.01, 0, 1.5708, 0, 0, d.origins + 12 * 7);
    originFrame(-0.045, 0, 0, 0, 1.5708, 0, d.origins + 12 * 8);
    originFrame(0, 0, 0, 0, 0, 0, d.origins + 12 * 9);  // gripper: the link frame itself
    originFrame(0.053125, 0, 0.603001417713939, 0, 0, 0, d.origins + 12 * 10);
    originFrame(0.053125, 0, 0.603001417713939 + 0.115 / 2, 0, 0, 0, d.origins + 12 * 11);
    // fetch_robot.hpp:38-53
    d.shapes[0] = {2, 0.29, (0.32 - 0.02) / 2, 0};
    d.shapes[1] = {0, (0.18 + 0.08) / 2, 0.34 / 2, 1.5 / 2};
    d.shapes[2] = {1, 0.0525, 0.29 / 2, 0};
    d.shapes[3] = {2, 0.06, 0.145 / 2, 0};
    d.shapes[4] = {1, 0.0525, 0.24 / 2, 0};
    d.shapes[5] = {2, 0.0525, 0.14 / 2, 0};
    d.shapes[6] = {1, 0.0525, 0.215 / 2, 0};
    d.shapes[7] = {2, 0.0525, 0.155 / 2, 0};
    d.shapes[8] = {1, 0.0525, 0.09 / 2, 0};
    d.shapes[9] = {0, (0.12 + 0.0625 * 2) / 2, 0.125 / 2, 0.075 / 2};
    d.shapes[10] = {1, 0.085 * 1.1, (0.04 * 1.1) / 2, 0};
    d.shapes[11] = {2, 0.24 * 1.05, (0.115 * 1.05) / 2, 0};
    {
        // FetchSceneDev::lift / reach.  Chain lengths from the shoulder pan joint's origin to each arm link's origin
        // (norms of the joint offsets in fetchFkKernel), plus the geometry's origin offset and bounding radius.
        const double chain[8] = {0, 0.13149, 0.35049, 0.48349, 0.68049, 0.80499, 0.94349, 0.94349 + 0.0765};   // pan .. gripper link
        for (int k = 0; k < 12; ++k) {
            d.lift[k] = k == 0 ? 0.f : 1.f;
            d.reach[k] = 0.f;
            if (k < 2 || k > 9) continue;   // base, torso, neck, head: the arm joints do not move them
            const FetchShape &sh = d.shapes[k];
            const double rad = sh.type == 0 ? std::sqrt(sh.a * sh.a + sh.b * sh.b + sh.c * sh.c)
                                            : sh.type == 1 ? sh.a + sh.b : std::sqrt(sh.a * sh.a + sh.b * sh.b);
            const double *o = d.origins + 12 * k;
            const double off = std::sqrt(o[3] * o[3] + o[7] * o[7] + o[11] * o[11]);
            d.reach[k] = (float)((chain[k - 2] + off + rad) * 1.0001);
        }
    }
    sc->envNodeCount = env.nodes.size();
    sc->robotNodeCount = 12;
    {
        // first-use costs (kernel load, call context, work lists) belong to construction
        const double q[8] = {0.1, 0, 0, 0, 0, 0, 0, 0};
        uint8_t v;
        if ((rc = sc->checkStates(q, 1, &v)) || (rc = sc->checkEdges(q, q, 1, &v, nullptr))) return rc;
        sc->hChecks = sc->hBv = sc->hTri = sc->hExact = sc->hLaunches = 0;
    }
    *out = reinterpret_cast<mplb_scene *>(static_cast<SceneBase *>(sc.release()));
    return MPLB_OK;
}

}  // namespace mplb

using namespace mplb;

extern "C" {

int mplb_scene_create_fetch(const double *env_tris, size_t n_env, const double env_frame[12], double check_resolution,
                            int device, mplb_scene **out) {
    try {
        return createFetch(env_tris, n_env, env_frame, check_resolution, device, out);
    } catch (const std::exception &e) {
        return fail(MPLB_ERR_INVALID, "%s", e.what());
    }
}

int mplb_scene_create_fetch_file(const char *env_path, const double env_frame[12], double check_resolution, int device,
                                 mplb_scene **out) {
    if (!env_path) return fail(MPLB_ERR_INVALID, "null path");
    try {
        // fetch_scenario.hpp:54: MeshLoad<Mesh>::load(envMesh, false, true)
        TriangleSoup env = loadColladaCached(env_path, false, true);
        return createFetch(env.tris.data(), env.size(), env_frame, check_resolution, device, out);
    } catch (const std::invalid_argument &e) {
        return fail(MPLB_ERR_IO, "%s", e.what());
    } catch (const std::exception &e) {
        return fail(MPLB_ERR_INVALID, "%s", e.what());
    }
}

}  // extern "C"
