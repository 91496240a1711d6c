// K1 batched-graph CSR builder, layer-0 fused aggregate+transform, K2 normalised-adjacency
// SpMM, K5 readout.  Integer work is bit-exact with the reference's dgl.batch / GraphConv
// degree normalisation; aggregation sums neighbours in ascending edge-id order
// with separate multiply and add roundings, i.e. the order torch's index_add_ uses on the CPU.
#include "common.cuh"
#include "launchers.h"

namespace eims {

// ------------------------------------------------------------------------------------ K1a
// One CTA: exclusive scan of atoms / directed edges per molecule -> gptr, eptr, dims.
__global__ void __launch_bounds__(1024) k1_scan_kernel(const int64_t* __restrict__ node_ptr,
                                                       const int64_t* __restrict__ bond_ptr,
                                                       const int32_t* __restrict__ ids, int B, int max_nodes,
                                                       int max_edges, int* __restrict__ gptr, int* __restrict__ eptr,
                                                       int* __restrict__ rowptr, int* __restrict__ dims) {
  pdl_sync();
  __shared__ int wn[32], we[32];
  __shared__ int carry_n, carry_e;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) { carry_n = 0; carry_e = 0; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    int g = base + tid, n = 0, e = 0;
    if (g < B) {
      int64_t id = ids ? (int64_t)ids[g] : (int64_t)g;
      n = (int)(node_ptr[id + 1] - node_ptr[id]);
      e = 2 * (int)(bond_ptr[id + 1] - bond_ptr[id]);
    }
    int in = n, ie = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int tn = __shfl_up_sync(0xffffffffu, in, o), te = __shfl_up_sync(0xffffffffu, ie, o);
      if (lane >= o) { in += tn; ie += te; }
    }
    if (lane == 31) { wn[w] = in; we[w] = ie; }
    __syncthreads();
    if (w == 0) {
      int a = wn[lane], b = we[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
        if (lane >= o) { a += ta; b += tb; }
      }
      wn[lane] = a; we[lane] = b;  // inclusive over warps
    }
    __syncthreads();
    int off_n = carry_n + (w ? wn[w - 1] : 0) + in - n;
    int off_e = carry_e + (w ? we[w - 1] : 0) + ie - e;
    if (g < B) { gptr[g] = off_n; eptr[g] = off_e; }
    __syncthreads();
    if (tid == 0) { carry_n += wn[31]; carry_e += we[31]; }
    __syncthreads();
  }
  if (tid == 0) {
    int N = carry_n, E = carry_e;
    bool over = N > max_nodes || E > max_edges || N < 0 || E < 0;
    gptr[B] = N; eptr[B] = E;
    dims[DIM_B] = over ? 0 : B;
    dims[DIM_N] = over ? 0 : N;
    dims[DIM_E] = over ? 0 : E;
    dims[DIM_ZERO_DEG] = 0;
    dims[DIM_OVERFLOW] = over ? 1 : 0;
    dims[5] = dims[6] = dims[7] = 0;
    rowptr[over ? 0 : N] = over ? 0 : E;
  }
}

// ------------------------------------------------------------------------------------ K1b
// One warp per molecule: features, COO edge list in reference order, CSR by destination
// (ascending edge id inside a row), per-atom graph id and degree normalisation.
__global__ void __launch_bounds__(256) k1_build_kernel(const int64_t* __restrict__ node_ptr,
                                                       const int64_t* __restrict__ bond_ptr,
                                                       const float* __restrict__ feat,
                                                       const int32_t* __restrict__ bond_begin,
                                                       const int32_t* __restrict__ bond_end,
                                                       const int32_t* __restrict__ ids, int B, int F,
                                                       const int* __restrict__ gptr, const int* __restrict__ eptr,
                                                       int* __restrict__ gid, int* __restrict__ src,
                                                       int* __restrict__ dst, int* __restrict__ rowptr,
                                                       int* __restrict__ col, float* __restrict__ norm,
                                                       float* __restrict__ x, int* __restrict__ dims) {
  pdl_sync();
  if (dims[DIM_OVERFLOW]) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int g = warp; g < B; g += nwarps) {
    const int64_t id = ids ? (int64_t)ids[g] : (int64_t)g;
    const int64_t a0 = node_ptr[id], b0 = bond_ptr[id];
    const int n = (int)(node_ptr[id + 1] - a0), nb = (int)(bond_ptr[id + 1] - b0);
    const int o = gptr[g], eo = eptr[g];
    for (int t = lane; t < n * F; t += 32) x[(int64_t)o * F + t] = __ldg(feat + a0 * F + t);
    const int32_t* bb = bond_begin + b0;
    const int32_t* be = bond_end + b0;
    for (int k = lane; k < nb; k += 32) {
      int b = __ldg(bb + k), e = __ldg(be + k);
      src[eo + 2 * k] = o + b; dst[eo + 2 * k] = o + e;          // GCN:142-143: [b->e, e->b]
      src[eo + 2 * k + 1] = o + e; dst[eo + 2 * k + 1] = o + b;
    }
    int run = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      int deg = 0;
      if (i < n)
        for (int k = 0; k < nb; ++k) deg += (__ldg(be + k) == i) + (__ldg(bb + k) == i);
      int inc = deg;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += t;
      }
      const int total = __shfl_sync(0xffffffffu, inc, 31);
      if (i < n) {
        int wpos = eo + run + inc - deg;
        rowptr[o + i] = wpos;
        gid[o + i] = g;
        // torch.pow(deg.float().clamp(min=1), -0.5) on the CPU == fl(1/fl(sqrt(d)))
        norm[o + i] = __fdiv_rn(1.0f, __fsqrt_rn((float)max(deg, 1)));
        if (deg == 0) atomicOr(dims + DIM_ZERO_DEG, 1);
        for (int k = 0; k < nb; ++k) {
          int b = __ldg(bb + k), e = __ldg(be + k);
          if (e == i) col[wpos++] = o + b;  // edge 2k   : b -> e
          if (b == i) col[wpos++] = o + e;  // edge 2k+1 : e -> b
        }
      }
      run += total;
    }
  }
}

int launch_csr_build(const eims_dataset* ds, const int32_t* ids, int B, int F, int max_nodes, int max_edges,
                     int* gptr, int* eptr, int* gid, int* src, int* dst, int* rowptr, int* col, float* norm,
                     float* x, int* dims, cudaStream_t st) {
  launch_pdl(k1_scan_kernel, dim3(1), dim3(1024), 0, st, ds->node_ptr, ds->bond_ptr, ids, B, max_nodes, max_edges, gptr, eptr, rowptr, dims);
  if (B > 0) {
    int blocks = (B + 7) / 8;
    launch_pdl(k1_build_kernel, dim3(blocks), dim3(256), 0, st, ds->node_ptr, ds->bond_ptr, ds->feat, ds->bond_begin, ds->bond_end, ids, B,
                                            F, gptr, eptr, gid, src, dst, rowptr, col, norm, x, dims);
  }
  return 0;
}

// --------------------------------------------------------------------------- layer 0 forward
// GraphConv(6 -> H) + ReLU in one pass (GCN:359-360 for i = 0): a0 = A (x*c) is 6 wide, so the
// transform is done in registers.  Warp per atom; writes a0 (saved for dW0) and z0 = relu(r).
constexpr int kMaxF0 = 8;

__global__ void __launch_bounds__(256) layer0_fwd_kernel(const int* __restrict__ dims, const int* __restrict__ rowptr,
                                                         const int* __restrict__ col, const float* __restrict__ norm,
                                                         const float* __restrict__ x, int F, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int H, float* __restrict__ a0,
                                                         float* __restrict__ z) {
  pdl_sync();
  const int N = dims[DIM_N];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < N; i += nwarps) {
    float a[kMaxF0];
#pragma unroll
    for (int f = 0; f < kMaxF0; ++f) a[f] = 0.f;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    for (int e = e0; e < e1; ++e) {
      const int j = __ldg(col + e);
      const float cj = __ldg(norm + j);
#pragma unroll
      for (int f = 0; f < kMaxF0; ++f)
        if (f < F) a[f] = __fadd_rn(a[f], __fmul_rn(__ldg(x + (int64_t)j * F + f), cj));
    }
    if (lane < F) {
      float v = 0.f;
#pragma unroll
      for (int f = 0; f < kMaxF0; ++f)
        if (f == lane) v = a[f];
      a0[(int64_t)i * F + lane] = v;
    }
    const float ci = __ldg(norm + i);
    for (int c0 = 0; c0 < H; c0 += 128) {
      const int c = c0 + 4 * lane;
      if (c < H) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < kMaxF0; ++f)
          if (f < F) {
            float4 w = ldg4(W + (int64_t)f * H + c);
            r.x = fmaf(a[f], w.x, r.x); r.y = fmaf(a[f], w.y, r.y);
            r.z = fmaf(a[f], w.z, r.z); r.w = fmaf(a[f], w.w, r.w);
          }
        float4 b = ldg4(bias + c);
        r.x = fmaxf(fmaf(r.x, ci, b.x), 0.f); r.y = fmaxf(fmaf(r.y, ci, b.y), 0.f);
        r.z = fmaxf(fmaf(r.z, ci, b.z), 0.f); r.w = fmaxf(fmaf(r.w, ci, b.w), 0.f);
        st4(z + (int64_t)i * H + c, r);
      }
    }
  }
}

int launch_layer0_fwd(const int* dims, const int* rowptr, const int* col, const float* norm, const float* x, int F,
                      const float* W, const float* bias, int H, float* a0, float* z, int max_nodes, cudaStream_t st) {
  if (F > kMaxF0) return EIMS_ERR_ARG;
  int blocks = (max_nodes + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(layer0_fwd_kernel, dim3(blocks), dim3(256), 0, st, dims, rowptr, col, norm, x, F, W, bias, H, a0, z);
  return 0;
}

// --------------------------------------------------------------------------- layer 0 weight grad
// dW0[f,k] += sum_i a0[i,f] * q[i,k]   (F <= 8 rows): block = 64 atoms, thread = 4 columns.
__global__ void __launch_bounds__(256) layer0_wgrad_kernel(const int* __restrict__ dims, const float* __restrict__ a0,
                                                           int F, const float* __restrict__ q, int H,
                                                           float* __restrict__ dW) {
  pdl_sync();
  const int N = dims[DIM_N];
  __shared__ float sa[64 * kMaxF0];
  const int cols4 = H >> 2;
  for (int r0 = blockIdx.x * 64; r0 < N; r0 += gridDim.x * 64) {
    const int rows = min(64, N - r0);
    __syncthreads();
    for (int t = threadIdx.x; t < rows * F; t += blockDim.x) sa[t] = a0[(int64_t)r0 * F + t];
    __syncthreads();
    for (int cg = threadIdx.x; cg < cols4; cg += blockDim.x) {
      float4 acc[kMaxF0];
#pragma unroll
      for (int f = 0; f < kMaxF0; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < rows; ++r) {
        float4 v = ldg4(q + (int64_t)(r0 + r) * H + 4 * cg);
#pragma unroll
        for (int f = 0; f < kMaxF0; ++f)
          if (f < F) {
            float a = sa[r * F + f];
            acc[f].x = fmaf(a, v.x, acc[f].x); acc[f].y = fmaf(a, v.y, acc[f].y);
            acc[f].z = fmaf(a, v.z, acc[f].z); acc[f].w = fmaf(a, v.w, acc[f].w);
          }
      }
#pragma unroll
      for (int f = 0; f < kMaxF0; ++f)
        if (f < F) {
          float* d = dW + (int64_t)f * H + 4 * cg;
          atomicAdd(d + 0, acc[f].x); atomicAdd(d + 1, acc[f].y);
          atomicAdd(d + 2, acc[f].z); atomicAdd(d + 3, acc[f].w);
        }
    }
  }
}

int launch_layer0_wgrad(const int* dims, const float* a0, int F, const float* q, int H, float* dW, int max_nodes,
                        cudaStream_t st) {
  if (F > kMaxF0) return EIMS_ERR_ARG;
  int blocks = (max_nodes + 63) / 64;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  launch_pdl(layer0_wgrad_kernel, dim3(blocks), dim3(256), 0, st, dims, a0, F, q, H, dW);
  return 0;
}

// ------------------------------------------------------------------------------------ K2
// Warp per destination row; lane owns NV float4 column groups (columns (i*32+lane)*4), so a
// whole row of H <= 128*NV floats is gathered by one warp-wide 128-bit load per neighbour and
// all NV*UE loads of a trip are independent (UE neighbours per trip).  Neighbours are added in
// ascending edge id, each with separate multiply and add roundings (torch's CPU index_add_).
//   forward  (out_mode 0): out_i = sum_j fl( drop(bn(h_j)) * c_j )
//   backward (out_mode 1): out_i = ( sum_j h_j ) * c_i * dropmask_i        (A symmetric)
template <int NV, int UE>
__global__ void __launch_bounds__(256) spmm_norm_kernel(const int* __restrict__ dims, const int* __restrict__ rowptr,
                                                        const int* __restrict__ col, const float* __restrict__ norm,
                                                        const float* __restrict__ h, int H,
                                                        const float* __restrict__ bn_scale,
                                                        const float* __restrict__ bn_shift, DropCfg drop, int out_mode,
                                                        float* __restrict__ out, int parts) {
  pdl_sync();
  const int N = dims[DIM_N];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool has_bn = bn_scale != nullptr;
  const bool in_drop = drop.active() && out_mode == 0;
  const bool out_drop = drop.active() && out_mode == 1;
  // `parts` warps share a row when H > 128*NV (warp = row*parts + part; part fixed per warp
  // because nwarps is a multiple of parts)
  const int cbase = (warp % parts) * (128 * NV);
  float4 sc[NV], sh[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = cbase + (v * 32 + lane) * 4;
    sc[v] = make_float4(1.f, 1.f, 1.f, 1.f);
    sh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bn && c < H) { sc[v] = ldg4(bn_scale + c); sh[v] = ldg4(bn_shift + c); }
  }
  for (int i = warp / parts; i < N; i += nwarps / parts) {
    const int e0 = __ldg(rowptr + i), e1 = __ldg(rowptr + i + 1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; e += UE) {
      int j[UE];
      float cj[UE];
      float4 val[UE][NV];
#pragma unroll
      for (int u = 0; u < UE; ++u) j[u] = (e + u < e1) ? __ldg(col + e + u) : -1;
#pragma unroll
      for (int u = 0; u < UE; ++u) {
        cj[u] = (out_mode == 0 && j[u] >= 0) ? __ldg(norm + j[u]) : 1.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = cbase + (v * 32 + lane) * 4;
          val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j[u] >= 0 && c < H) val[u][v] = ldg4(h + (int64_t)j[u] * H + c);
        }
      }
#pragma unroll
      for (int u = 0; u < UE; ++u) {
        if (j[u] < 0) continue;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = cbase + (v * 32 + lane) * 4;
          if (c >= H) continue;
          float4 x = val[u][v];
          if (has_bn) {
            x.x = fmaf(x.x, sc[v].x, sh[v].x); x.y = fmaf(x.y, sc[v].y, sh[v].y);
            x.z = fmaf(x.z, sc[v].z, sh[v].z); x.w = fmaf(x.w, sc[v].w, sh[v].w);
          }
          if (in_drop) {
            const float4 m = drop_mask4(drop, (uint64_t)j[u] * H + c);
            x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
          }
          if (out_mode == 0) {
            x.x = __fmul_rn(x.x, cj[u]); x.y = __fmul_rn(x.y, cj[u]);
            x.z = __fmul_rn(x.z, cj[u]); x.w = __fmul_rn(x.w, cj[u]);
          }
          acc[v].x = __fadd_rn(acc[v].x, x.x); acc[v].y = __fadd_rn(acc[v].y, x.y);
          acc[v].z = __fadd_rn(acc[v].z, x.z); acc[v].w = __fadd_rn(acc[v].w, x.w);
        }
      }
    }
    const float ci = out_mode == 1 ? __ldg(norm + i) : 1.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = cbase + (v * 32 + lane) * 4;
      if (c >= H) continue;
      float4 a = acc[v];
      if (out_mode == 1) {
        a.x *= ci; a.y *= ci; a.z *= ci; a.w *= ci;
        if (out_drop) {
          const float4 m = drop_mask4(drop, (uint64_t)i * H + c);
          a.x *= m.x; a.y *= m.y; a.z *= m.z; a.w *= m.w;
        }
      }
      st4(out + (int64_t)i * H + c, a);
    }
  }
}

int launch_spmm_norm(const int* dims, const int* rowptr, const int* col, const float* norm, const float* h, int H,
                     const float* bn_scale, const float* bn_shift, DropCfg drop, int out_mode, float* out,
                     int max_nodes, cudaStream_t st) {
  if (H % 4) return EIMS_ERR_ARG;
  const int parts = H <= 512 ? 1 : (H + 511) / 512;
  if (parts > 8 || (8 % parts)) return EIMS_ERR_ARG;  // 8 warps per block must split evenly over a row
  int blocks = (max_nodes * parts + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
#define EIMS_SPMM(NV, UE) \
  launch_pdl(spmm_norm_kernel<NV, UE>, dim3(blocks), dim3(256), 0, st, dims, rowptr, col, norm, h, H, bn_scale, bn_shift, drop, out_mode, out, parts)
  if (H <= 128) EIMS_SPMM(1, 4);
  else if (H <= 256) EIMS_SPMM(2, 2);
  else EIMS_SPMM(4, 1);
#undef EIMS_SPMM
  return 0;
}

// ------------------------------------------------------------------------------------ K5
// Block per graph: BN apply + segment sum / mean / max / sum||max with the first arg-max
// (DGL's CPU SegmentCmp keeps the first maximum in node order).  256 threads = H/4 float4
// column lanes x (1024/H) row lanes; row lane k takes nodes r0+k, r0+k+RL, ... and the row
// lanes are combined in shared memory (ties between lanes go to the smaller node id).
__global__ void __launch_bounds__(256) readout_kernel(const int* __restrict__ dims, const int* __restrict__ gptr,
                                                      const float* __restrict__ z, int H,
                                                      const float* __restrict__ bn_scale,
                                                      const float* __restrict__ bn_shift, int pooling,
                                                      float* __restrict__ out, int* __restrict__ argmax) {
  pdl_sync();
  __shared__ float4 ssum[256], smax[256];
  __shared__ int4 sarg[256];
  const int B = dims[DIM_B];
  const int g = blockIdx.x;
  if (g >= B) return;
  const int cpl = H >> 2;                       // float4 columns per row
  const int CL = cpl < 256 ? cpl : 256;         // column lanes
  const int RL = 256 / CL;                      // row lanes
  const int cl = threadIdx.x % CL, rl = threadIdx.x / CL;
  const int pool_dim = pooling == EIMS_POOL_COMBINED ? 2 * H : H;
  const bool has_bn = bn_scale != nullptr;
  const int r0 = __ldg(gptr + g), r1 = __ldg(gptr + g + 1);
  const float ninf = -__int_as_float(0x7f800000);
  for (int cb = 0; cb < cpl; cb += CL) {
    const int c = (cb + cl) * 4;
    const bool on = rl < RL && cb + cl < cpl;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(ninf, ninf, ninf, ninf);
    int4 am = make_int4(r0, r0, r0, r0);
    if (on) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bn) { sc = ldg4(bn_scale + c); sh = ldg4(bn_shift + c); }
#pragma unroll 4
      for (int i = r0 + rl; i < r1; i += RL) {
        float4 v = ldg4(z + (int64_t)i * H + c);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        if (v.x > mx.x) { mx.x = v.x; am.x = i; }
        if (v.y > mx.y) { mx.y = v.y; am.y = i; }
        if (v.z > mx.z) { mx.z = v.z; am.z = i; }
        if (v.w > mx.w) { mx.w = v.w; am.w = i; }
      }
    }
    __syncthreads();
    ssum[threadIdx.x] = s; smax[threadIdx.x] = mx; sarg[threadIdx.x] = am;
    __syncthreads();
    if (rl == 0 && on) {
      for (int k = 1; k < RL; ++k) {
        const float4 s2 = ssum[k * CL + cl], m2 = smax[k * CL + cl];
        const int4 a2 = sarg[k * CL + cl];
        s.x += s2.x; s.y += s2.y; s.z += s2.z; s.w += s2.w;
        if (m2.x > mx.x || (m2.x == mx.x && a2.x < am.x)) { mx.x = m2.x; am.x = a2.x; }
        if (m2.y > mx.y || (m2.y == mx.y && a2.y < am.y)) { mx.y = m2.y; am.y = a2.y; }
        if (m2.z > mx.z || (m2.z == mx.z && a2.z < am.z)) { mx.z = m2.z; am.z = a2.z; }
        if (m2.w > mx.w || (m2.w == mx.w && a2.w < am.w)) { mx.w = m2.w; am.w = a2.w; }
      }
      float* o = out + (int64_t)g * pool_dim;
      if (pooling == EIMS_POOL_MEAN) {
        const float n = (float)(r1 - r0);  // torch: S / n
        s.x = s.x / n; s.y = s.y / n; s.z = s.z / n; s.w = s.w / n;
      }
      if (pooling == EIMS_POOL_SUM || pooling == EIMS_POOL_MEAN || pooling == EIMS_POOL_COMBINED) st4(o + c, s);
      if (pooling == EIMS_POOL_MAX) st4(o + c, mx);
      if (pooling == EIMS_POOL_COMBINED) st4(o + H + c, mx);
      if (argmax && (pooling == EIMS_POOL_MAX || pooling == EIMS_POOL_COMBINED))
        *reinterpret_cast<int4*>(argmax + (int64_t)g * H + c) = am;
    }
  }
}

int launch_readout(const int* dims, const int* gptr, const float* z, int H, const float* bn_scale,
                   const float* bn_shift, int pooling, float* out, int* argmax, int max_graphs, cudaStream_t st) {
  if (H % 4 || H < 4) return EIMS_ERR_ARG;
  launch_pdl(readout_kernel, dim3(max_graphs < 1 ? 1 : max_graphs), dim3(256), 0, st, dims, gptr, z, H, bn_scale, bn_shift, pooling, out, argmax);
  return 0;
}

}  // namespace eims
