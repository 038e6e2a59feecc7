// b200audio — generated-mel (T, M) float32 log-mel kernels with the epilogue inside the mel phase and a TMA write-out
// (included by fast_fwd.cuh, inside namespace b2a::<anonymous>; uses its fill / stage-1 / stage-2 device functions).
//
// Common to both kernels (n_mels % 32 == 0, max-type guard — the Whisper family):
//   * the warp-uniform constant tables (window, inter-stage twiddles) live in TENSOR MEMORY (tm_* in fast_fwd.cuh);
//   * the epilogue (guard, MUFU log2, folded affine, running max / min) runs inside the mel phase, lane == frame;
//   * the finished tile leaves through the TMA: the staging tile is laid out as the boxes of the output tensor map (32
//     frames x 32 features, 128-byte swizzle — the XOR that makes the lane-strided STS.128 of the mel phase bank-conflict
//     free) and ONE thread issues ONE bulk tensor store per tile (cp.async.bulk.tensor.4d, SASS UTMASTG; 4-D map: 32 features
//     x 32 frames x n_mels / 32 boxes x 1 clip).  No write-out phase: no LDS / STG of the staging tile by the warps (-128
//     shared-memory wavefronts and ~900 warp-instructions per tile); rows past the last frame are clipped by the map.
//
//   fast_logmel_tma_kernel      the production kernel: one warp group, C::MIN_BLOCKS CTAs per SM, three barriers per tile.
//   fast_logmel_ws_tma_kernel   the same work as a WARP-SPECIALISED two-tile pipeline (B2A_WS=1): one CTA of 2 * C::WARPS
//                               warps per SM, producer warps [0, WARPS) run fill + stage 1 of tile i while consumer warps
//                               [WARPS, 2 * WARPS) run stage 2 + the mel phase of tile i - 1; exchange buffer and sample
//                               tile double buffered; the groups meet only at named barriers (bar.arrive on one side,
//                               bar.sync on the other — the producer / consumer pattern of the PTX manual):
//                                 FULL[b]   P -> C   exchange buffer b holds stage 1 of a tile
//                                 EMPTY[b]  C -> P   the consumers have consumed exchange buffer b (power tile in place)
//                                 P, C               group-local barriers
//                               Measured (4096 x 30 s, Whisper-128): 5.15 ms against 5.13 ms for the single-group kernel —
//                               the groups overlap stage 1 (shared-memory pipe) with stage 2 (FP32 pipe) by construction, but
//                               the consumers' chain (stage 2 + mel) is 40 % longer than the producers' and the producers idle
//                               a third of the time; dealing the mel rows to both groups (2 * WARPS parts) balanced the
//                               chains but pushed the warp-specialised straight-line code past the 32 KB instruction cache
//                               (hit rate 75 %, 7.0 ms).  DESIGN.md section 6 has the whole series.
#pragma once

// TMA variant: one dense staging tile of M / 32 boxes (32 rows x 128 bytes each), 1024-byte aligned — the dynamic
// shared-memory window itself is declared 1024-byte aligned
template <class C, int M>
struct SmemWST {
  static constexpr int cdiv4(int bytes) { return (bytes + 15) / 16; }
  static constexpr int EX = 0;
  static constexpr int EX_SIZE = cdiv4(8 * C::FT * C::EP);
  static constexpr int YT = (EX + 2 * EX_SIZE + 63) / 64 * 64;
  static constexpr int YT_SIZE = cdiv4(4 * C::FT * M);
  static constexpr int XS = YT + YT_SIZE;
  static constexpr int XS_SIZE = cdiv4(4 * C::XS_FLOATS);
  static constexpr int END = XS + 2 * XS_SIZE;
};

enum : int { WSB_FULL = 1, WSB_EMPTY = 3, WSB_P = 5, WSB_C = 6 };
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// the tile walk both groups share: g = blockIdx.x + i * gridDim.x, i = 0 .. n_my - 1, as (clip, tile) without a division per step
struct WsWalk {
  int tpc, step_c, step_t, n_my, clip_i, tile_i;
  __device__ __forceinline__ void init(const FastParams& p) {
    tpc = p.tiles_per_clip;
    const int64_t total = (int64_t)p.batch * tpc;
    n_my = (int64_t)blockIdx.x < total ? (int)((total - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    step_c = (int)(gridDim.x / (unsigned)tpc);
    step_t = (int)(gridDim.x - (unsigned)step_c * (unsigned)tpc);
    clip_i = (int)(blockIdx.x / (unsigned)tpc);
    tile_i = (int)(blockIdx.x - (unsigned)clip_i * (unsigned)tpc);
  }
  __device__ __forceinline__ void advance(int& c, int& t) const {
    c += step_c;
    t += step_t;
    if (t >= tpc) {
      t -= tpc;
      ++c;
    }
  }
  __device__ __forceinline__ void back(int& c, int& t) const {
    c -= step_c;
    t -= step_t;
    if (t < 0) {
      t += tpc;
      --c;
    }
  }
};

// every warp parks its constant tables in tensor memory: window + inter-stage twiddles of its stage-1 role (producers), the
// post-twiddles of its stage-2 unit (consumers); 4 * N1 columns per warp slot, five slots per lane quadrant
template <class C>
__device__ __forceinline__ uint32_t ws_park_tables(const FastParams& p, uint32_t tm_base, int warp) {
  static_assert(C::RPW == 1 && 4 * C::N1 >= 2 * C::N2 && 5 * 4 * C::N1 <= 512 && 2 * C::WARPS <= 20, "tensor-memory table layout");
  const uint32_t tmc = tm_warp_base(tm_base, warp, 4 * C::N1);
  if (warp < C::WARPS) {
    tm_store_table(tmc, p.win2 + warp * C::N1, C::N1);
    tm_store_table(tmc + 2 * C::N1, p.tw1 + warp * C::N1, C::N1);
  } else {
    const int u = (C::WARPS - 1) - (warp - C::WARPS);  // stage2_tile's unit of this consumer warp
    tm_store_table(tmc, p.twp + u * 2 * C::N2, C::N2);
  }
  return tmc;
}

// GUARD_MAX: the guard is max(a, floor) (Whisper family) — no add in the epilogue; otherwise max(a + add, floor)
template <class C, class MS, int SPECK, bool GUARD_MAX>
__global__ void __launch_bounds__(2 * C::THREADS, 1) fast_logmel_ws_tma_kernel(const FastParams p, const __grid_constant__ CUtensorMap out_map) {
  static_assert(MS::M > 0 && MS::NW == C::WARPS && MS::F == C::F && MS::M % 32 == 0, "mel spec / kernel variant mismatch");
  constexpr int NT = 2 * C::THREADS, NBOX = MS::M / 32;
  using S = SmemWST<C, MS::M>;
  extern __shared__ __align__(1024) float4 smem_ws4[];
  float4* const smem4 = smem_ws4;
  float2* const E0 = reinterpret_cast<float2*>(smem4 + S::EX);
  constexpr unsigned EX_FLOAT2 = 2u * S::EX_SIZE, XS_BYTES = 16u * S::XS_SIZE;
  const unsigned y_sa0 = (unsigned)__cvta_generic_to_shared(smem4 + S::YT);
  __shared__ float red_max[1][C::WARPS], red_min[1][C::WARPS];
  __shared__ uint32_t s_tm_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TM_COLS = 512;
  const uint32_t tm_base = tm_alloc<TM_COLS>(&s_tm_slot, warp);
  const uint32_t tmc = ws_park_tables<C>(p, tm_base, warp);

  WsWalk wk;
  wk.init(p);
  const int n_my = wk.n_my;
  int clip_i = wk.clip_i, tile_i = wk.tile_i;
  const bool want_max = p.clip_max != nullptr;

  if (warp < C::WARPS) {
    // ================================ producer group: fill, stage 1 =================================================
    float* const xs0 = reinterpret_cast<float*>(smem4 + S::XS) + C::XS_HEAD;
    FillCtx<C> fc;
    fc.template init<0>(p, (unsigned)__cvta_generic_to_shared(xs0));
    if (n_my > 0) fill_tile<C, 0>(p, xs0, fc, clip_i, tile_i, 0u);
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
      const int b = i & 1;
      int nclip = clip_i, ntile = tile_i;
      wk.advance(nclip, ntile);
      cp_async_wait_all();
      named_bar_sync(WSB_P, C::THREADS);  // sample tile b landed; every producer warp is past stage 1 of tile i - 1
      if (i + 1 < n_my)
        fill_tile<C, 0>(p, reinterpret_cast<float*>(reinterpret_cast<char*>(xs0) + (b ^ 1) * XS_BYTES), fc, nclip, ntile, (b ^ 1) * XS_BYTES);
      if (i >= 2) named_bar_sync(WSB_EMPTY + b, NT);
      stage1_tile<C, 0, true>(reinterpret_cast<const float*>(reinterpret_cast<const char*>(xs0) + b * XS_BYTES), E0 + b * EX_FLOAT2, nullptr,
                              nullptr, warp, lane, 0.0f, tmc);
      named_bar_arrive(WSB_FULL + b, NT);
      clip_i = nclip;
      tile_i = ntile;
    }
  } else {
    // ================================ consumer group: stage 2, mel phase with the epilogue, TMA stores ================
    const int cw = warp - C::WARPS;
    const bool pw_only = SPECK >= 0 ? SPECK == B2A_SPEC_POWER : p.spec_kind == B2A_SPEC_POWER;
    const float spec_eps = p.spec_eps;
    const float guard_add = p.guard_add, guard_floor = p.guard_floor, y_mul = p.y_mul, y_add = p.y_add;
    const bool use_log = p.use_log != 0;
    const unsigned x7 = (unsigned)(lane & 7) << 4;
    // The group's first warp sends the tile off: ONE bulk tensor store moves the whole staging tile (4-D map: 32 features x 32
    // frames x n_mels / 32 boxes x 1 clip); the per-warp max / min are folded by lanes.
    auto send_tile = [&](int jclip, int jtile) {
      if (cw != 0) return;
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&out_map), "r"(0), "r"(jtile * C::FT),
                     "r"(0), "r"(jclip), "r"(y_sa0)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (want_max) {
        const float a = lane < C::WARPS ? red_max[0][lane] : -INFINITY, m = lane < C::WARPS ? red_min[0][lane] : INFINITY;
        const int kmax = __reduce_max_sync(0xffffffffu, float_key(a));
        const int kmin = __reduce_min_sync(0xffffffffu, float_key(m));
        if (lane == 0) {
          atomic_max_f(p.clip_max + jclip, key_float(kmax));
          p.tile_min[(int64_t)jclip * p.tile_min_pitch + jtile] = key_float(kmin);
        }
      }
    };
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
      const int b = i & 1;
      float2* const E = E0 + b * EX_FLOAT2;
      named_bar_sync(WSB_FULL + b, NT);
      stage2_tile<C, true, false, true>(E, nullptr, nullptr, cw, lane, pw_only, spec_eps, tmc);
      if (threadIdx.x == C::THREADS) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the store of tile i - 1 has read the staging tile
      named_bar_sync(WSB_C, C::THREADS);  // power tile complete (in place in the exchange buffer)
      // ---- mel phase (lane == frame): banded projection from the in-place power tile, epilogue, running max / min; the row
      // quad goes to row `lane` of box m / 32 of the staging tile — 16-byte chunk (m % 32) / 4 lands at chunk ^ (row & 7),
      // the tensor map's 128-byte swizzle
      float lmax = -INFINITY, lmin = INFINITY;
      {
        const float* pr = reinterpret_cast<const float*>(E + lane * C::EP) + (lane >> 4);
        const unsigned yrow_sa = y_sa0 + 128u * (unsigned)lane;
        MS::template run<C>(cw, pr, [&](auto M_, float a0, float a1, float a2, float a3) {
          constexpr int m = decltype(M_)::value;
          float e[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float a = GUARD_MAX ? fmaxf(e[c], guard_floor) : fmaxf(e[c] + guard_add, guard_floor);
            const float y = use_log ? lg2_approx(a) : a;
            e[c] = fmaf(y, y_mul, y_add);
          }
          lmax = fmax3(lmax, e[0], e[1]);
          lmin = fmin3(lmin, e[0], e[1]);
          lmax = fmax3(lmax, e[2], e[3]);
          lmin = fmin3(lmin, e[2], e[3]);
          const unsigned addr = yrow_sa + ((unsigned)(((m % 32) / 4) << 4) ^ x7);
          asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"((m / 32) * 4096), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]));
        });
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staging tile is read by the async proxy
      if (i + 2 < n_my) named_bar_arrive(WSB_EMPTY + b, NT);        // exchange buffer b is free for stage 1 of tile i + 2
      if (want_max) {
        const int64_t left = p.frame_count - (int64_t)tile_i * C::FT;
        if (lane >= left) {  // frames past the end of the clip: computed on zero samples, clipped by the tensor map
          lmax = -INFINITY;
          lmin = INFINITY;
        }
        const int kmax = __reduce_max_sync(0xffffffffu, float_key(lmax));
        const int kmin = __reduce_min_sync(0xffffffffu, float_key(lmin));
        if (lane == 0) {
          red_max[0][cw] = key_float(kmax);
          red_min[0][cw] = key_float(kmin);
        }
      }
      named_bar_sync(WSB_C, C::THREADS);  // staging tile complete
      send_tile(clip_i, tile_i);
      wk.advance(clip_i, tile_i);
    }
    if (threadIdx.x == C::THREADS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last tile has left shared memory
  }
  tm_free<TM_COLS>(tm_base, warp);
}

// ---- fast_logmel_tma_kernel: the single-group kernel (C::MIN_BLOCKS CTAs per SM, three barriers per tile) with the same
// epilogue-in-the-mel-phase + TMA write-out: stage-1 tables in tensor memory, post-twiddles in shared memory.
template <class C, int M>
struct SmemT {
  static constexpr int cdiv4(int bytes) { return (bytes + 15) / 16; }
  static constexpr int YT = 0;  // staging tile first: 1024-byte aligned like the window itself
  static constexpr int YT_SIZE = cdiv4(4 * C::FT * M);
  static constexpr int EX = YT + YT_SIZE;
  static constexpr int EX_SIZE = cdiv4(8 * C::FT * C::EP);
  static constexpr int TWP = EX + EX_SIZE;
  static constexpr int XS = TWP + cdiv4(8 * C::NC);
  static constexpr int END = XS + cdiv4(4 * C::XS_FLOATS);
};

// FUSED (cooperative launch): the whole clamping forward in ONE launch.  No initialised statistics: the tile's maximum goes to
// tile_max[clip][tile] with a plain store; after the last tile a grid-wide barrier, then the clamp fix-up by the same CTAs —
// a CTA takes whole clips: reduces the clip's tile maxima, and its warps rewrite (or, for unwritten all-silent tiles, write) the
// tiles whose minimum lies below the floor.  Three launches -> one: at 512 clips per GPU (BASELINE's batch on 8 GPUs) the two
// small launches were 3 % of the step, for a single 30 s clip a third of the call.
template <class C, int M>
__device__ __forceinline__ void fused_clamp_fixup(const FastParams& p, float c_zero, float* red, int warp, int lane) {
  const int tpc = p.tiles_per_clip;
  for (int clip = blockIdx.x; clip < p.batch; clip += gridDim.x) {
    float m = -INFINITY;
    for (int t = threadIdx.x; t < tpc; t += C::THREADS) m = fmaxf(m, __ldcg(p.tile_max + (int64_t)clip * p.tile_min_pitch + t));
    m = key_float(__reduce_max_sync(0xffffffffu, float_key(m)));
    __syncthreads();  // red is free (previous clip's readers are done)
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < C::WARPS; ++w) m = fmaxf(m, red[w]);
    if (p.clip_max && threadIdx.x == 0) p.clip_max[clip] = m;  // caller-visible statistic
    const float floor_out = m - p.clamp_delta;
    for (int t = warp; t < tpc; t += C::WARPS) {
      const float tmin = __ldcg(p.tile_min + (int64_t)clip * p.tile_min_pitch + t);
      const bool unwritten = tmin == -INFINITY;
      if (!unwritten && !(tmin < floor_out)) continue;
      const int64_t left = p.frame_count - (int64_t)t * C::FT;
      const int n4 = (int)(left < C::FT ? left : C::FT) * (M / 4);
      float4* t4 = reinterpret_cast<float4*>(p.out + (int64_t)clip * p.out_clip_stride + (int64_t)t * C::FT * M);
      if (unwritten) {
        const float v = fmaxf(c_zero, floor_out);
        const float4 v4 = make_float4(v, v, v, v);
        for (int i = lane; i < n4; i += 32) t4[i] = v4;
      } else {
        for (int i = lane; i < n4; i += 32) {
          float4 v = __ldcg(t4 + i);
          if (v.x < floor_out || v.y < floor_out || v.z < floor_out || v.w < floor_out) {  // (a NaN stays a NaN)
            v.x = v.x < floor_out ? floor_out : v.x;
            v.y = v.y < floor_out ? floor_out : v.y;
            v.z = v.z < floor_out ? floor_out : v.z;
            v.w = v.w < floor_out ? floor_out : v.w;
            t4[i] = v;
          }
        }
      }
    }
  }
}

template <class C, class MS, int SPECK, bool GUARD_MAX, bool FUSED = false>
__global__ void __launch_bounds__(C::THREADS, C::MIN_BLOCKS) fast_logmel_tma_kernel(const FastParams p, const __grid_constant__ CUtensorMap out_map) {
  static_assert(MS::M > 0 && MS::NW == C::WARPS && MS::F == C::F && MS::M % 32 == 0 && C::TM_OK, "mel spec / kernel variant mismatch");
  constexpr int NC = C::NC;
  using S = SmemT<C, MS::M>;
  extern __shared__ __align__(1024) float4 smem_ws4[];
  float4* const smem4 = smem_ws4;
  float2* const s_twp = reinterpret_cast<float2*>(smem4 + S::TWP);
  float2* const E = reinterpret_cast<float2*>(smem4 + S::EX);
  float* const xs = reinterpret_cast<float*>(smem4 + S::XS) + C::XS_HEAD;
  const unsigned y_sa = (unsigned)__cvta_generic_to_shared(smem4 + S::YT);
  __shared__ float red_max[C::WARPS], red_min[C::WARPS];
  __shared__ uint32_t s_tm_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NC; i += C::THREADS) s_twp[i] = p.twp[i];
  const uint32_t tm_base = tm_alloc<C::TM_COLS>(&s_tm_slot, warp);
  const uint32_t tmc = tm_warp_base(tm_base, warp, C::TM_PER_WARP);
  for (int rr = 0; rr < C::RPW; ++rr) {
    const int n2 = warp * C::RPW + rr;
    tm_store_table(tmc + rr * 4 * C::N1, p.win2 + n2 * C::N1, C::N1);
    tm_store_table(tmc + rr * 4 * C::N1 + 2 * C::N1, p.tw1 + n2 * C::N1, C::N1);
  }
  FillCtx<C> fc;
  fc.template init<0>(p, (unsigned)__cvta_generic_to_shared(xs));

  WsWalk wk;
  wk.init(p);
  const int n_my = wk.n_my;
  int clip_i = wk.clip_i, tile_i = wk.tile_i;
  const bool want_max = FUSED || p.clip_max != nullptr;
  const bool pw_only = SPECK >= 0 ? SPECK == B2A_SPEC_POWER : p.spec_kind == B2A_SPEC_POWER;
  const float spec_eps = p.spec_eps;
  const float guard_add = p.guard_add, guard_floor = p.guard_floor, y_mul = p.y_mul, y_add = p.y_add;
  const bool use_log = p.use_log != 0;
  const unsigned x7 = (unsigned)(lane & 7) << 4;
  // the finished tile: ONE bulk tensor store (4-D map: 32 features x 32 frames x n_mels / 32 boxes x 1 clip) by warp 0, which also
  // folds the per-warp max / min (lanes)
  // A tile that holds nothing but the guard-floor constant c (every frame digitally silent: c is the smallest value the
  // epilogue can produce, so "tile max <= c" means every element equals c) will be overwritten by the clamp anyway unless
  // the whole clip is silent: with p.skip_floor_tiles it is not stored at all, its tile_min entry becomes -inf, and the
  // fix-up WRITES max(c, floor) there — one write instead of write + read + write (30 % silence: +11 % -> see DESIGN).
  const float c_floor = fmaf(use_log ? lg2_approx(fmaxf(0.0f + guard_add, guard_floor)) : fmaxf(0.0f + guard_add, guard_floor), y_mul, y_add);
  const bool skip_ok = want_max && p.skip_floor_tiles != 0;
  auto send_tile = [&](int jclip, int jtile) {
    if (warp != 0) return;
    bool store = true;
    if (want_max) {
      const float a = lane < C::WARPS ? red_max[lane] : -INFINITY, m = lane < C::WARPS ? red_min[lane] : INFINITY;
      const int kmax = __reduce_max_sync(0xffffffffu, float_key(a));
      const int kmin = __reduce_min_sync(0xffffffffu, float_key(m));
      const float tmax = key_float(kmax);
      store = !(skip_ok && tmax <= c_floor);
      if (lane == 0) {
        if constexpr (FUSED) p.tile_max[(int64_t)jclip * p.tile_min_pitch + jtile] = tmax;
        else atomic_max_f(p.clip_max + jclip, tmax);
        p.tile_min[(int64_t)jclip * p.tile_min_pitch + jtile] = store ? key_float(kmin) : -INFINITY;
      }
    }
    if (lane == 0 && store) {
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&out_map), "r"(0), "r"(jtile * C::FT),
                   "r"(0), "r"(jclip), "r"(y_sa)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  };
  if (n_my > 0) fill_tile<C, 0>(p, xs, fc, clip_i, tile_i);
  int pclip = 0, ptile = 0;
#pragma unroll 1
  for (int i = 0; i < n_my; ++i) {
    int nclip = clip_i, ntile = tile_i;
    wk.advance(nclip, ntile);
    cp_async_wait_all();
    __syncthreads();  // sample tile landed; the mel phase of tile i - 1 is complete (staging tile, per-warp max / min)
    if (i > 0) send_tile(pclip, ptile);
    stage1_tile<C, 0, true>(xs, E, nullptr, nullptr, warp, lane, 0.0f, tmc);
    __syncthreads();  // exchange buffer complete, sample tile free
    if (i + 1 < n_my) fill_tile<C, 0>(p, xs, fc, nclip, ntile);  // the next tile's samples land during stage 2 / mel
    stage2_tile<C, true>(E, nullptr, s_twp, warp, lane, pw_only, spec_eps);
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the store of tile i - 1 has read the staging tile
    __syncthreads();  // power tile complete (in place in the exchange buffer)
    // ---- mel phase (lane == frame): banded projection, epilogue, running max / min; row quad -> row `lane` of box m / 32 of
    // the staging tile, 16-byte chunk (m % 32) / 4 at chunk ^ (row & 7) (the tensor map's 128-byte swizzle)
    float lmax = -INFINITY, lmin = INFINITY;
    {
      const float* pr = reinterpret_cast<const float*>(E + lane * C::EP) + (lane >> 4);
      const unsigned yrow_sa = y_sa + 128u * (unsigned)lane;
      MS::template run<C>(warp, pr, [&](auto M_, float a0, float a1, float a2, float a3) {
        constexpr int m = decltype(M_)::value;
        float e[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float a = GUARD_MAX ? fmaxf(e[c], guard_floor) : fmaxf(e[c] + guard_add, guard_floor);
          const float y = use_log ? lg2_approx(a) : a;
          e[c] = fmaf(y, y_mul, y_add);
        }
        lmax = fmax3(lmax, e[0], e[1]);
        lmin = fmin3(lmin, e[0], e[1]);
        lmax = fmax3(lmax, e[2], e[3]);
        lmin = fmin3(lmin, e[2], e[3]);
        const unsigned addr = yrow_sa + ((unsigned)(((m % 32) / 4) << 4) ^ x7);
        asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"((m / 32) * 4096), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]));
      });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staging tile is read by the async proxy
    if (want_max) {
      const int64_t left = p.frame_count - (int64_t)tile_i * C::FT;
      if (lane >= left) {  // frames past the end of the clip: computed on zero samples, clipped by the tensor map
        lmax = -INFINITY;
        lmin = INFINITY;
      }
      const int kmax = __reduce_max_sync(0xffffffffu, float_key(lmax));
      const int kmin = __reduce_min_sync(0xffffffffu, float_key(lmin));
      if (lane == 0) {
        red_max[warp] = key_float(kmax);
        red_min[warp] = key_float(kmin);
      }
    }
    pclip = clip_i;
    ptile = tile_i;
    clip_i = nclip;
    tile_i = ntile;
  }
  cp_async_wait_all();
  __syncthreads();
  if (n_my > 0) send_tile(pclip, ptile);
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last tile has left shared memory
  if constexpr (FUSED) {
    // every CTA's tiles and statistics are complete (bulk stores waited for by their issuing threads) and fenced: grid barrier,
    // then the clamp over whole clips
    __threadfence();
    cooperative_groups::this_grid().sync();
    fused_clamp_fixup<C, MS::M>(p, c_floor, red_max, warp, lane);
  }
  tm_free<C::TM_COLS>(tm_base, warp);
}
