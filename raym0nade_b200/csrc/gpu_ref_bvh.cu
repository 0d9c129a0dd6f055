// gpu_ref_bvh.cu — BVH::build (src/bvh.cpp:18-54) on the device, and the refit of every tree of an uploaded scene for moved
// vertices ("device-side BVH flatten/refit and GPU build that reproduces the reference topology").
//
// The reference's rule: a node over faces [L, R) with more than 10 faces splits at M = (L + R) / 2 after std::nth_element along
// the axis whose centroid variance is largest (axis 0 unless D[1] > D[0]; axis 2 if D[2] beats both); children sit at 2u and
// 2u + 1 of a heap-indexed array; a leaf stores its range and the union of its faces' boxes, an inner node the union of its
// children's.  Because M depends on L and R only, every node's range follows from n and u alone - so a whole level is built at
// once:
//   1. variances    per face slot the centre, in double; an inclusive scan (cub::DeviceScan) turns them into prefix sums, so the mean
//                   of every segment of the level is two reads; a second scan over the squared deviations from the slot's own
//                   segment mean gives n * variance per segment, accurate even for an eleven-face node far from the origin
//   2. axis         per node of the level: the reference's comparison rule on those (its D = Em2 - Em * Em / n)
//   3. keys + sort  per face slot: (segment start << 32) | order-preserving bits of the centroid along the segment's axis; one
//                   radix sort of the frame (cub::DeviceRadixSort) sorts every segment of the level at once - a sorted
//                   range is one of the orders std::nth_element may leave; slots in leaves of earlier levels keep their place
// and after the last level one pass writes the leaves (ranges + boxes, the reference's min / max order) and one pass per level,
// bottom up, the inner boxes.  ~2 ms for a million triangles against ~150 ms for the host's nth_element recursion.
//
// What "reproduces the reference topology" means: the same rule, hence the same tree shape (node count, ranges, heap layout)
// and - whenever no two variances tie within rounding and no two centroids tie across a median - the same face SETS in every
// node.  Bit-identity of the permutation is not defined by the reference: the order std::nth_element leaves inside a range is
// an artefact of libstdc++'s introselect, and the fp32 running sums the reference forms over that order decide near-ties of
// the axis choice.  (The variances here are formed in double.)  The boxes of a node depend on its face set alone and are the
// reference's bits for that set.  tests/test_gpu_trace.py::test_device_built_reference_tree.
//
// Refit (rm_scene_refit): vertices moved, topology kept.  The reference tree's boxes are recomputed by the same two passes; the
// 4-wide quantised tree(s) bottom-up through parent links - a node is re-quantised by the last of its children to finish
// (one atomic counter per node), from exact child boxes kept in a side array so that quantisation slack does not compound.
#include <algorithm>
#include <cmath>
#include <functional>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "rm_context.cuh"
#include "wide_bvh.h"
#include "gpu_bvh.h"

namespace {

constexpr int kLeafBag = 10;                 // LeafBagSize, src/bvh.cpp:16

struct Moments { double s[3]; };
struct MomentsSum {
    __host__ __device__ Moments operator()(const Moments &a, const Moments &b) const {
        Moments r;
#pragma unroll
        for (int k = 0; k < 3; k++) r.s[k] = a.s[k] + b.s[k];
        return r;
    }
};

// Range [L, R) of heap node u over n faces; false when the node does not exist (an ancestor already was a leaf).
__host__ __device__ inline bool node_range(int u, int n, int &L, int &R) {
    int d = 0;
    while ((u >> (d + 1)) != 0) d++;
    L = 0; R = n;
    for (int b = d - 1; b >= 0; b--) {
        if (R - L <= kLeafBag) return false;
        const int M = (L + R) / 2;
        if ((u >> b) & 1) L = M; else R = M;
    }
    return true;
}

// The node of level `level` (or the leaf above it) that holds face slot i: its range; returns its heap index.
__device__ __forceinline__ int slot_node(int i, int n, int level, int &L, int &R) {
    int u = 1;
    L = 0; R = n;
    for (int d = 0; d < level; d++) {
        if (R - L <= kLeafBag) break;
        const int M = (L + R) / 2;
        if (i >= M) { L = M; u = u << 1 | 1; } else { R = M; u = u << 1; }
    }
    return u;
}

// Face::center() (src/component.cpp:37-39): (v0 + v1 + v2) / 3.0f, glm's vec3 / scalar = multiplication by 1.0f / 3.0f
__device__ __forceinline__ float centre_axis(const float *p, int a) { return __fmul_rn(__fadd_rn(__fadd_rn(p[a], p[3 + a]), p[6 + a]), __frcp_rn(3.0f)); }
__device__ __forceinline__ double centre_finite(const float *p, int a) {
    const double c = double(centre_axis(p, a));
    return isfinite(c) ? c : 0.0;                   // a non-finite triangle can never be hit; it must not poison a whole segment
}

__global__ void k_ref_init(int n, int *order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[i] = i;
}

// The variance of a segment comes in two steps, each a prefix sum over the whole frame of slots, so that it is exact to double
// rounding of SMALL numbers even for an eleven-face node of a large scene far from the origin: first the segment means (prefix
// sums of the centres), then the sums of squared deviations from the slot's own segment mean.
__global__ void k_ref_centres(const float *__restrict__ pos, const int *__restrict__ order, int n, Moments *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pos + size_t(order[i]) * 9;
    Moments m;
#pragma unroll
    for (int a = 0; a < 3; a++) m.s[a] = centre_finite(p, a);
    out[i] = m;
}

__global__ void k_ref_means(const Moments *__restrict__ prefix, int n, int level, Moments *__restrict__ mean) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (1 << level)) return;
    int L, R;
    Moments m = {{0.0, 0.0, 0.0}};
    if (node_range((1 << level) + k, n, L, R) && R - L > kLeafBag) {
        const Moments hi = prefix[R - 1];
#pragma unroll
        for (int a = 0; a < 3; a++) m.s[a] = (hi.s[a] - (L ? prefix[L - 1].s[a] : 0.0)) / double(R - L);
    }
    mean[k] = m;
}

__global__ void k_ref_deviations(const float *__restrict__ pos, const int *__restrict__ order, int n, int level, const Moments *__restrict__ mean,
                                 Moments *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int L, R;
    const int u = slot_node(i, n, level, L, R);
    Moments m = {{0.0, 0.0, 0.0}};
    if (u >= (1 << level) && R - L > kLeafBag) {
        const float *p = pos + size_t(order[i]) * 9;
        const Moments mu = mean[u - (1 << level)];
#pragma unroll
        for (int a = 0; a < 3; a++) { const double d = centre_finite(p, a) - mu.s[a]; m.s[a] = d * d; }
    }
    out[i] = m;
}

__global__ void k_ref_axis(const Moments *__restrict__ prefix, int n, int level, signed char *__restrict__ axis) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (1 << level)) return;
    int L, R;
    signed char ax = -1;
    if (node_range((1 << level) + k, n, L, R) && R - L > kLeafBag) {
        const Moments hi = prefix[R - 1];
        double D[3];                                 // n * variance per axis; the reference's D = Em2 - Em * Em / n (src/bvh.cpp:26-33)
#pragma unroll
        for (int a = 0; a < 3; a++) D[a] = hi.s[a] - (L ? prefix[L - 1].s[a] : 0.0);
        ax = 0;
        if (D[1] > D[0]) ax = 1;
        if (D[2] > D[0] && D[2] > D[1]) ax = 2;
    }
    axis[k] = ax;
}

__global__ void k_ref_keys(const float *__restrict__ pos, const int *__restrict__ order, int n, int level, const signed char *__restrict__ axis,
                           unsigned long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int L, R;
    const int u = slot_node(i, n, level, L, R);
    unsigned low = unsigned(i - L);                  // a slot of a finished leaf keeps its place
    if (u >= (1 << level) && R - L > kLeafBag) {
        const int ax = axis[u - (1 << level)];
        const float c = centre_axis(pos + size_t(order[i]) * 9, ax);
        unsigned b = __float_as_uint(c);
        if (c != c) b = 0xffffffffu;                 // NaN centroids go last
        else b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        low = b;
    }
    keys[i] = (static_cast<unsigned long long>(unsigned(L)) << 32) | low;
}

// Leaves of the tree being built: their ranges (the boxes follow below)
__global__ void k_ref_leaf_ranges(int n, int n_nodes, RmBvhNode *__restrict__ nodes) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < 1 || u >= n_nodes) return;
    int L, R;
    if (!node_range(u, n, L, R) || R - L > kLeafBag) return;
    nodes[u].faceL = L;
    nodes[u].faceR = R;
}

// Leaf boxes of any heap-indexed tree (leaf iff faceR != 0, src/bvh.cpp:57).  Face::aabb is glm::min(v0, glm::min(v1, v2)) /
// glm::max alike - glm::min(x, y) = (y < x) ? y : x - and Box + Box is std::fmin / std::fmax per component
// (include/geometry.h), folded over the faces in order from (+INF, -INF).  order == nullptr: the positions already are in tree
// order (refit).
__global__ void k_ref_leaf_boxes(const float *__restrict__ pos, const int *__restrict__ order, int n, int n_nodes, RmBvhNode *__restrict__ nodes) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < 1 || u >= n_nodes) return;
    const int L = nodes[u].faceL, R = nodes[u].faceR;
    if (R == 0 || L < 0 || L >= R || R > n) return;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = L; i < R; i++) {
        const float *p = pos + size_t(order ? order[i] : i) * 9;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float x = p[a], y = p[3 + a], z = p[6 + a];
            const float mn_yz = (z < y) ? z : y, mn = (mn_yz < x) ? mn_yz : x;
            const float mx_yz = (y < z) ? z : y, mx = (x < mx_yz) ? mx_yz : x;
            lo[a] = fminf(lo[a], mn);
            hi[a] = fmaxf(hi[a], mx);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) { nodes[u].v0[a] = lo[a]; nodes[u].v1[a] = hi[a]; }
}

// Inner boxes of one level: the union of the children's (src/bvh.cpp:41).  A never-written slot (all zero) has never-written
// children and stays all zero.
__global__ void k_ref_inner(int level, int n_nodes, RmBvhNode *__restrict__ nodes) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (1 << level)) return;
    const int u = (1 << level) + k;
    if ((u << 1 | 1) >= n_nodes || nodes[u].faceR != 0) return;
    const RmBvhNode a = nodes[u << 1], b = nodes[u << 1 | 1];
#pragma unroll
    for (int c = 0; c < 3; c++) { nodes[u].v0[c] = fminf(a.v0[c], b.v0[c]); nodes[u].v1[c] = fmaxf(a.v1[c], b.v1[c]); }
}

int tree_levels(int n) {          // levels that hold an inner node: the rightmost path carries the largest ranges (src/bvh.cpp:44-46)
    int levels = 0;
    for (int s = n; s > kLeafBag; s = (s + 1) >> 1) levels++;
    return levels;
}

int launch_boxes(const float *d_pos, const int *d_order, int n, int n_nodes, RmBvhNode *d_nodes, cudaStream_t st, uint64_t &launches) {
    k_ref_leaf_boxes<<<(n_nodes + 255) / 256, 256, 0, st>>>(d_pos, d_order, n, n_nodes, d_nodes);
    launches++;
    int top = 0;                                     // level of the last slot
    while ((int64_t(2) << top) <= int64_t(n_nodes - 1)) top++;
    for (int level = top - 1; level >= 0; level--) {
        k_ref_inner<<<((1 << level) + 255) / 256, 256, 0, st>>>(level, n_nodes, d_nodes);
        launches++;
    }
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

// ---- refit of the 4-wide tree
struct WideFit {
    RmWideNode *nodes;
    const int *order;          // tree order -> face
    const float *pos;          // [n][9], face order
    int *parent, *pending;
    float4 *ex_lo, *ex_hi;     // exact box per record
    int *bad;
};

__global__ void k_wide_links(WideFit W, int n_nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (i == 0) W.parent[0] = -1;
    const RmWideNode &w = W.nodes[i];
    int inner = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (w.meta[k] & 0x80) { W.parent[w.child_base + (w.meta[k] & 0x7f)] = i; inner++; }
    W.pending[i] = inner;
}

__device__ void wide_refit_node(const WideFit &W, int i) {
    RmWideNode w = W.nodes[i];
    float lo[4][3], hi[4][3];
    int n = 0;
    for (int k = 0; k < 4; k++) {
        const unsigned m = w.meta[k];
        if (!m) break;                               // children fill the slots from 0
        if (m & 0x80) {
            const int c = w.child_base + int(m & 0x7f);
            const float4 l = __ldcg(W.ex_lo + c), h = __ldcg(W.ex_hi + c);          // written by another thread: read past L1
            lo[k][0] = l.x; lo[k][1] = l.y; lo[k][2] = l.z; hi[k][0] = h.x; hi[k][1] = h.y; hi[k][2] = h.z;
        } else {
            const int first = w.tri_base + int(m >> 2), cnt = int(m & 3);
            for (int a = 0; a < 3; a++) { lo[k][a] = CUDART_INF_F; hi[k][a] = -CUDART_INF_F; }
            for (int t = 0; t < cnt; t++) {
                const float *p = W.pos + size_t(W.order[first + t]) * 9;
                for (int a = 0; a < 3; a++) {
                    lo[k][a] = fminf(lo[k][a], fminf(fminf(p[a], p[3 + a]), p[6 + a]));
                    hi[k][a] = fmaxf(hi[k][a], fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]));
                }
            }
        }
        n = k + 1;
    }
    if (n == 0) return;
    if (!wide_quantise(lo, hi, n, w)) atomicExch(W.bad, 1);
    float4 ulo = make_float4(lo[0][0], lo[0][1], lo[0][2], 0.0f), uhi = make_float4(hi[0][0], hi[0][1], hi[0][2], 0.0f);
    for (int k = 1; k < n; k++) {
        ulo.x = fminf(ulo.x, lo[k][0]); ulo.y = fminf(ulo.y, lo[k][1]); ulo.z = fminf(ulo.z, lo[k][2]);
        uhi.x = fmaxf(uhi.x, hi[k][0]); uhi.y = fmaxf(uhi.y, hi[k][1]); uhi.z = fmaxf(uhi.z, hi[k][2]);
    }
    __stcg(W.ex_lo + i, ulo);
    __stcg(W.ex_hi + i, uhi);
    const uint4 *src = reinterpret_cast<const uint4 *>(&w);
    uint4 *dst = reinterpret_cast<uint4 *>(W.nodes + i);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];          // o, s, qlo, qhi; the links (bytes 48..63) are unchanged
}

__global__ void k_wide_refit(WideFit W, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || W.pending[i] != 0) return;   // start at the records whose children are all leaves ...
    for (;;) {
        wide_refit_node(W, i);
        const int p = W.parent[i];
        if (p < 0) return;
        __threadfence();                             // ... and climb: the last child to arrive re-quantises the parent
        if (atomicSub(W.pending + p, 1) != 1) return;
        __threadfence();
        i = p;
    }
}

} // namespace

// declared in host_prep.cpp
int rm_prepare_scene_impl(const RmRawScene *raw, RmPrepared **out,
                          const std::function<int(const float *, int, std::vector<RmBvhNode> &, std::vector<int32_t> &)> *tree);

static int build_on_device(RmContext *ctx, const float *positions, int n, RmBvhNode *nodes, int n_nodes, int32_t *perm) {
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf *B = ctx->b_build;
    int rc;
    // [0] keys a, [1] keys b, [2] order a, [3] order b, [4] per-slot terms, [5] their prefix sums, [6] positions, [7] nodes, [8] segment means,
    // [12] axis, [15] cub temp
    if ((rc = B[0].alloc(size_t(n) * 8)) || (rc = B[1].alloc(size_t(n) * 8)) || (rc = B[2].alloc(size_t(n) * 4)) || (rc = B[3].alloc(size_t(n) * 4)) ||
        (rc = B[4].alloc(size_t(n) * sizeof(Moments))) || (rc = B[5].alloc(size_t(n) * sizeof(Moments))) || (rc = B[6].alloc(size_t(n) * 36)) ||
        (rc = B[7].alloc(size_t(n_nodes) * sizeof(RmBvhNode))))
        return rc;
    const int levels = tree_levels(n);
    if ((rc = B[12].alloc(size_t(1) << std::max(levels, 1))) || (rc = B[8].alloc((size_t(1) << std::max(levels, 1)) * sizeof(Moments)))) return rc;
    int end_bit = 33;
    while (end_bit < 64 && (uint64_t(n) >> (end_bit - 32)) != 0) end_bit++;
    size_t temp_sort = 0, temp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, B[0].as<unsigned long long>(), B[1].as<unsigned long long>(), B[2].as<int>(), B[3].as<int>(), n, 0, end_bit, st);
    cub::DeviceScan::InclusiveScan(nullptr, temp_scan, B[4].as<Moments>(), B[5].as<Moments>(), MomentsSum(), n, st);
    const size_t temp_bytes = std::max(temp_sort, temp_scan);
    if ((rc = B[15].alloc(temp_bytes))) return rc;

    const float *d_pos = B[6].as<float>();
    RM_CUDA(cudaMemcpyAsync(B[6].p, positions, size_t(n) * 36, cudaMemcpyHostToDevice, st));
    RM_CUDA(cudaMemsetAsync(B[7].p, 0, size_t(n_nodes) * sizeof(RmBvhNode), st));          // never-written slots stay zero
    const int grid_n = (n + 255) / 256;
    int *order = B[2].as<int>(), *order_alt = B[3].as<int>();
    k_ref_init<<<grid_n, 256, 0, st>>>(n, order);
    ctx->launches++;
    for (int level = 0; level < levels; level++) {
        size_t tb = temp_bytes;
        k_ref_centres<<<grid_n, 256, 0, st>>>(d_pos, order, n, B[4].as<Moments>());
        RM_CUDA(cub::DeviceScan::InclusiveScan(B[15].p, tb, B[4].as<Moments>(), B[5].as<Moments>(), MomentsSum(), n, st));
        k_ref_means<<<((1 << level) + 255) / 256, 256, 0, st>>>(B[5].as<Moments>(), n, level, B[8].as<Moments>());
        k_ref_deviations<<<grid_n, 256, 0, st>>>(d_pos, order, n, level, B[8].as<Moments>(), B[4].as<Moments>());
        tb = temp_bytes;
        RM_CUDA(cub::DeviceScan::InclusiveScan(B[15].p, tb, B[4].as<Moments>(), B[5].as<Moments>(), MomentsSum(), n, st));
        k_ref_axis<<<((1 << level) + 255) / 256, 256, 0, st>>>(B[5].as<Moments>(), n, level, B[12].as<signed char>());
        k_ref_keys<<<grid_n, 256, 0, st>>>(d_pos, order, n, level, B[12].as<signed char>(), B[0].as<unsigned long long>());
        tb = temp_bytes;
        RM_CUDA(cub::DeviceRadixSort::SortPairs(B[15].p, tb, B[0].as<unsigned long long>(), B[1].as<unsigned long long>(), order, order_alt, n, 0, end_bit, st));
        std::swap(order, order_alt);
        ctx->launches += 6;
    }
    k_ref_leaf_ranges<<<(n_nodes + 255) / 256, 256, 0, st>>>(n, n_nodes, B[7].as<RmBvhNode>());
    ctx->launches++;
    if ((rc = launch_boxes(d_pos, order, n, n_nodes, B[7].as<RmBvhNode>(), st, ctx->launches))) return rc;
    RM_CUDA(cudaMemcpyAsync(nodes, B[7].p, size_t(n_nodes) * sizeof(RmBvhNode), cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaMemcpyAsync(perm, order, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

static int refit_wide(RmContext *ctx, DevBuf &b_nodes, DevBuf &b_order, int n_nodes) {
    if (n_nodes <= 0) return RM_OK;
    cudaStream_t st = ctx->stream;
    DevBuf *B = ctx->b_build;
    int rc;
    if ((rc = B[8].alloc(size_t(n_nodes) * 4)) || (rc = B[9].alloc(size_t(n_nodes) * 4)) || (rc = B[4].alloc(size_t(n_nodes) * 16)) ||
        (rc = B[5].alloc(size_t(n_nodes) * 16)) || (rc = B[14].alloc(64)))
        return rc;
    RM_CUDA(cudaMemsetAsync(B[14].p, 0, 64, st));
    WideFit W{b_nodes.as<RmWideNode>(), b_order.as<int>(), ctx->b_raw[0].as<float>(), B[8].as<int>(), B[9].as<int>(), B[4].as<float4>(), B[5].as<float4>(), B[14].as<int>()};
    const int grid = (n_nodes + 127) / 128;
    k_wide_links<<<grid, 128, 0, st>>>(W, n_nodes);
    k_wide_refit<<<grid, 128, 0, st>>>(W, n_nodes);
    ctx->launches += 2;
    RM_CUDA(cudaGetLastError());
    int bad = 0;
    RM_CUDA(cudaMemcpyAsync(&bad, B[14].p, 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    if (bad) return rm_fail(RM_ERR_INVALID, "rm_scene_refit: a node could not enclose its children (non-finite positions?)");
    return RM_OK;
}


extern "C" {

int32_t rm_tree_node_count(int32_t n_faces) {
    if (n_faces <= 0) return 0;
    int u = 1;
    for (int s = n_faces; s > kLeafBag; s = (s + 1) >> 1) u = u << 1 | 1;          // nodeCount, src/bvh.cpp:44-46
    return u + 1;
}

int rm_tree_build(RmContext *ctx, const float *positions, int32_t n_faces, RmBvhNode *nodes, int32_t n_nodes, int32_t *perm) {
    if (!ctx || !positions || !nodes || !perm) return rm_fail(RM_ERR_INVALID, "rm_tree_build: null argument");
    if (n_faces <= 0) return rm_fail(RM_ERR_INVALID, "rm_tree_build: no faces");
    if (n_nodes != rm_tree_node_count(n_faces)) return rm_fail(RM_ERR_INVALID, "rm_tree_build: %d faces need %d nodes, not %d", n_faces, rm_tree_node_count(n_faces), n_nodes);
    return build_on_device(ctx, positions, n_faces, nodes, n_nodes, perm);
}

int rm_prepare_scene_device(RmContext *ctx, const RmRawScene *raw, RmPrepared **out) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "rm_prepare_scene_device: context is NULL");
    const std::function<int(const float *, int, std::vector<RmBvhNode> &, std::vector<int32_t> &)> tree =
        [ctx](const float *positions, int n, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &perm) {
            nodes.assign(size_t(rm_tree_node_count(n)), RmBvhNode{});
            perm.resize(size_t(n));
            return build_on_device(ctx, positions, n, nodes.data(), int(nodes.size()), perm.data());
        };
    return rm_prepare_scene_impl(raw, out, &tree);
}

int rm_scene_refit(RmContext *ctx, const float *positions, int32_t n_faces) {
    if (!ctx || !positions) return rm_fail(RM_ERR_INVALID, "rm_scene_refit: null argument");
    if (!ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_scene_refit: no scene uploaded");
    if (n_faces != ctx->scene.n_faces) return rm_fail(RM_ERR_INVALID, "rm_scene_refit: %d faces given, the uploaded scene has %d", n_faces, ctx->scene.n_faces);
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = ctx->scene.n_faces, n_nodes = ctx->scene.n_nodes;
    int rc;
    ctx->have_primary = ctx->have_gbuffer = ctx->have_resolved = false;
    // a background build still running was started for the old vertices: its tree is dropped when it ends
    if (ctx->refine) ctx->refine->discard = true;
    RM_CUDA(cudaMemcpyAsync(ctx->b_raw[0].p, positions, size_t(n) * 36, cudaMemcpyHostToDevice, st));
    RM_CUDA(cudaStreamSynchronize(st));              // pageable source
    ctx->scene_h2d_bytes = int64_t(n) * 36;
    if ((rc = launch_boxes(ctx->b_raw[0].as<float>(), nullptr, n, n_nodes, ctx->b_nodes.as<RmBvhNode>(), st, ctx->launches))) return rc;
    if (ctx->have_wide && (rc = refit_wide(ctx, ctx->b_nodes_wide, ctx->b_facemap_wide, ctx->wide_nodes))) return rc;
    if (ctx->have_wide && ctx->refined_installed && (rc = refit_wide(ctx, ctx->b_nodes_wide2, ctx->b_facemap_wide2, ctx->refined_nodes))) return rc;
    // trees keyed by the old geometry must not be picked up for it again
    ctx->refined_key = 0; ctx->refined_n = -1;
    ctx->fast_key_valid = false;
    return rm_repack_faces(ctx);
}

} // extern "C"
