// Host emulation of the per-thread device GJK (jg_gjk.cuh compiled with g++, JG_HOST_EMU).
// Test infrastructure: lets the CPU suite check the fp32 kernel numerics against the double oracle.
#define JG_HOST_EMU 1
#include "../../jogramop_framework_b200/csrc/jg_epa.cuh"

#include <algorithm>
#include <vector>

using namespace jg;

static int g_small_storage = 0;
static long g_overflows = 0;
extern "C" void emu_set_small_storage(int on) { g_small_storage = on; g_overflows = 0; }
extern "C" long emu_overflows() { return g_overflows; }

// va [na,3], vb [nb,3] float; T [n,12] row-major 3x4 (A placed by T, B static); out dist[n], status[n], iters[n]
extern "C" void emu_gjk_batch(const float* va, int na, const float* vb, int nb, const float* T, int n, float dmax,
                              float* dist, int* status, int* iters, float* depth, int* epa_iters) {
    // device layout: vertex sets padded to a multiple of 4 with copies of their last vertex
    std::vector<float4> A((na + 3) / 4 * 4), B((nb + 3) / 4 * 4);
    for (size_t i = 0; i < A.size(); ++i) { const int k = std::min<int>((int)i, na - 1); A[i] = make_float4(va[3 * k], va[3 * k + 1], va[3 * k + 2], 0.f); }
    for (size_t i = 0; i < B.size(); ++i) { const int k = std::min<int>((int)i, nb - 1); B[i] = make_float4(vb[3 * k], vb[3 * k + 1], vb[3 * k + 2], 0.f); }
    for (int k = 0; k < n; ++k) {
        const float* t = T + 12 * k;
        Tf X;
        X.r00 = t[0]; X.r01 = t[1]; X.r02 = t[2]; X.tx = t[3];
        X.r10 = t[4]; X.r11 = t[5]; X.r12 = t[6]; X.ty = t[7];
        X.r20 = t[8]; X.r21 = t[9]; X.r22 = t[10]; X.tz = t[11];
        float d = 0.f;
        int it = 0;
        Simplex S;
        status[k] = gjk_distance(A.data(), na, B.data(), nb, X, mk(X.tx, X.ty, X.tz), dmax, d, it, S);
        dist[k] = d;
        iters[k] = it;
        if (status[k] == JG_GJK_PENETRATING && depth) {
            int eit = 0;
            if (g_small_storage) {
                // the fast kernel's compact shared-memory storage class (14 expansions) with the overflow fallback
                uint32_t words[EpaCompact<1>::WORDS];
                EpaStateC<1> E;
                E.poly.w = words;
                int st = JG_EPA_DONE;
                const ThreadSupport sup = {A.data(), na, B.data(), nb, X};
                if (epa_begin(E, sup, S)) {
                    do { st = epa_step(E, sup); } while (st == JG_EPA_RUNNING);
                    depth[k] = epa_result(E);
                    eit = E.it;
                } else {
                    depth[k] = 0.f;
                }
                if (st == JG_EPA_OVERFLOW) {
                    g_overflows++;
                    depth[k] = epa_depth(A.data(), na, B.data(), nb, X, S, eit);
                }
            } else {
                depth[k] = epa_depth(A.data(), na, B.data(), nb, X, S, eit);
            }
            epa_iters[k] = eit;
        }
    }
}

extern "C" void emu_plane_batch(const float* va, int na, const float* T, int n, const float* plane, float* dist) {
    std::vector<float4> A(na);
    for (int i = 0; i < na; ++i) A[i] = make_float4(va[3 * i], va[3 * i + 1], va[3 * i + 2], 0.f);
    for (int k = 0; k < n; ++k) {
        const float* t = T + 12 * k;
        Tf X;
        X.r00 = t[0]; X.r01 = t[1]; X.r02 = t[2]; X.tx = t[3];
        X.r10 = t[4]; X.r11 = t[5]; X.r12 = t[6]; X.ty = t[7];
        X.r20 = t[8]; X.r21 = t[9]; X.r22 = t[10]; X.tz = t[11];
        dist[k] = plane_distance(A.data(), na, X, mk(plane[0], plane[1], plane[2]), plane[3]);
    }
}
