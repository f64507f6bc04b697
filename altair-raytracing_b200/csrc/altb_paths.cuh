// altb_paths.cuh -- per-ray polylines for small N: what ARay::MakePolyLine3D feeds the reference's OpenGL views
// (makeIntegratingSphereNRays.C:69-72, makeIntegratingSphere1Ray.C:21-53).  Point 0 = source, then every surface hit,
// then the world-box point of an exited ray.  One thread per ray, same bounce_step as the trace kernel.
#pragma once
#include "altb_kernels.cuh"

namespace altb {

template <bool ROUGH, int MODEL>
__global__ void __launch_bounds__(128) k_trace_paths(const __grid_constant__ TraceParams P, f3 src_pos, uint32_t max_points,
                                                     float* __restrict__ pts, uint32_t* __restrict__ npts,
                                                     uint8_t* __restrict__ status) {
    constexpr bool NEED_G = ROUGH || MODEL == 1;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const DrawTabs T = make_tabs(P.sincos);
    float* p = pts + (size_t)i * max_points * 3;
    uint32_t np_ = 0;
    auto put = [&](const f3& v) {
        if (np_ < max_points) { p[3 * np_] = v.x; p[3 * np_ + 1] = v.y; p[3 * np_ + 2] = v.z; }
        np_++;
    };
    put(src_pos);
    RayState s;
    s.pos = {(float)P.x0[0], (float)P.x0[1], (float)P.x0[2]};
    s.dir = {(float)P.d0[0], (float)P.d0[1], (float)P.d0[2]};
    s.hits = 0; s.where = EV_WALL;
    int st = 0;
    if (P.kind0 == EV_EXIT) st = ALTB_EXITED; else s.where = P.kind0;
    while (!st) {
        put(s.pos);
        HitDraws dr;
        const uint64_t rid = P.ray_id0 + i;
        hit_from_philox<NEED_G>(P.keys, T, P.k.abs_thr, P.k.spec_thr, (uint32_t)rid, (uint32_t)(rid >> 32), s.hits, dr);
        if (MODEL == 3) dr.u_r = lobe_accept(P.keys, rid, s.hits, P.k.lobe_n, P.k.lobe_ang);
        st = bounce_step<ROUGH, MODEL, false>(P.g, P.k, P.k.zc, s, dr);
    }
    if (st == ALTB_EXITED) put(s.pos);
    npts[i] = np_;
    if (status) status[i] = (uint8_t)st;
}

}  // namespace altb
