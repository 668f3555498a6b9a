// mpp_map.cu -- grid-map handle: env.py grid (list-of-lists of 0/1/2/3) -> border-padded,
// bit-packed occupancy tensor in HBM + the per-map safety-class table (helper.py:67-80).
#include <stdarg.h>
#include <stdlib.h>

#include <cmath>

#include "mpp_common.cuh"

static thread_local char g_err[512] = "";

void mpp_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int mpp_abi_version(void) { return MPP_ABI_VERSION; }
extern "C" const char *mpp_last_error(void) { return g_err; }

extern "C" int mpp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int mpp_check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        mpp_set_error("no CUDA device visible (%s); libmpp_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MPP_ENODEVICE;
    }
    if (device < 0 || device >= n) { mpp_set_error("device %d out of range [0,%d)", device, n); return MPP_EINVAL; }
    int major = 0;
    MPP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
        mpp_set_error("device %d is sm_%d0; libmpp_b200 is built for sm_100a only (no fallback)", device, major);
        return MPP_ENODEVICE;
    }
    return MPP_OK;
}

// One thread per padded word: word (pr, pw) holds padded columns [32*pw, 32*pw+32) of padded row pr.
__global__ void mpp_pack_occ_kernel(const uint8_t *__restrict__ grid, int rows, int cols, int pitch_words,
                                    int total_words, uint32_t *__restrict__ occ) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total_words) return;
    grid += (size_t)blockIdx.y * rows * cols;                  // blockIdx.y = map of a batch
    occ += (size_t)blockIdx.y * total_words;
    int pr = w / pitch_words, pw = w % pitch_words;
    uint32_t bits = 0;
    if (pr >= rows + 2) { occ[w] = 0xffffffffu; return; }  // tail padding
    int r = pr - 1;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
        int c = pw * 32 + b - 1;
        bool blocked = (r < 0 || r >= rows || c < 0 || c >= cols) ? true : (grid[(size_t)r * cols + c] == 1);
        bits |= (blocked ? 1u : 0u) << b;
    }
    occ[w] = bits;
}

// Static per-cell move mask in MAACO's move order (MAACO.py:98): bit m set iff the move stays in bounds,
// lands on a non-obstacle cell (MAACO.py:93-95 without the tabu test) and, for diagonals, does not cut an
// obstacle corner (MAACO.py:100-120).  0 for obstacle cells.
__global__ void mpp_svalid_kernel(const uint32_t *__restrict__ occ, int occ_words, int pitch, int R, int C,
                                  uint8_t *__restrict__ sv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    occ += (size_t)blockIdx.y * occ_words;                     // blockIdx.y = map of a batch
    sv += (size_t)blockIdx.y * R * C;
    const int r = i / C, c = i % C;
    auto blocked = [&](int rr, int cc) -> bool {
        const int pb = cc + 1;
        return (occ[(rr + 1) * pitch + (pb >> 5)] >> (pb & 31)) & 1u;
    };
    uint32_t mask = 0;
    if (!blocked(r, c)) {
        for (int m = 0; m < 8; ++m) {
            const int dr = (int)((0xA940u >> (2 * m)) & 3u) - 1, dc = (int)((0x9224u >> (2 * m)) & 3u) - 1;
            bool ok = !blocked(r + dr, c + dc);
            if (ok && dr != 0 && dc != 0) ok = !blocked(r + dr, c) && !blocked(r, c + dc);
            mask |= (ok ? 1u : 0u) << m;
        }
    }
    sv[i] = (uint8_t)mask;
}

extern "C" int mpp_map_create(const uint8_t *grid_host, int rows, int cols, int device, mpp_map **out) {
    MPP_REQUIRE(grid_host && out, "mpp_map_create: null argument");
    MPP_REQUIRE(rows > 0 && cols > 0 && (long long)rows * cols < (1ll << 30), "mpp_map_create: bad shape %dx%d", rows, cols);
    int rc = mpp_check_device(device);
    if (rc) return rc;
    MPP_CUDA(cudaSetDevice(device));
    mpp_map *m = (mpp_map *)calloc(1, sizeof(mpp_map));
    if (!m) { mpp_set_error("out of host memory"); return MPP_ENOMEM; }
    m->rows = rows; m->cols = cols; m->device = device;
    m->start = m->target = -1;
    m->safety_msd = -1.0;
    size_t n = (size_t)rows * cols;
    m->grid_host = (uint8_t *)malloc(n);
    memcpy(m->grid_host, grid_host, n);
    for (size_t i = 0; i < n; ++i) {
        uint8_t v = grid_host[i];
        if (v == 2 && m->start < 0) m->start = (int)i;   // first row-major hit, MAACO.py:32-41
        if (v == 3 && m->target < 0) m->target = (int)i;
        if (v == 1) m->n_obstacles++;
    }
    m->pitch_words = (cols + 2 + 31) / 32;
    int words = (rows + 2) * m->pitch_words;
    m->occ_words = (words + 3) & ~3;
    MPP_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
    uint8_t *gdev = nullptr;
    MPP_CUDA(cudaMalloc(&gdev, n));
    MPP_CUDA(cudaMalloc(&m->occ_dev, (size_t)m->occ_words * 4));
    MPP_CUDA(cudaMemcpy(gdev, grid_host, n, cudaMemcpyHostToDevice));
    mpp_pack_occ_kernel<<<(m->occ_words + 255) / 256, 256>>>(gdev, rows, cols, m->pitch_words, m->occ_words, m->occ_dev);
    MPP_CUDA(cudaGetLastError());
    MPP_CUDA(cudaMalloc(&m->svalid_dev, n));
    mpp_svalid_kernel<<<((int)n + 255) / 256, 256>>>(m->occ_dev, m->occ_words, m->pitch_words, rows, cols, m->svalid_dev);
    MPP_CUDA(cudaGetLastError());
    // the map as a batch of one (what the MAACO entry points take)
    mpp_map_batch &B = m->self_batch;
    B.n_maps = 1; B.rows = rows; B.cols = cols; B.device = device; B.sm_count = m->sm_count;
    B.pitch_words = m->pitch_words; B.occ_words = m->occ_words;
    B.occ_dev = m->occ_dev; B.svalid_dev = m->svalid_dev; B.grid_host = m->grid_host; B.owns = 0;
    B.meta_host = (MppMapMeta *)malloc(sizeof(MppMapMeta));
    *B.meta_host = mpp_make_meta(rows, cols, m->start, m->target);
    MPP_CUDA(cudaMalloc(&B.meta_dev, sizeof(MppMapMeta)));
    MPP_CUDA(cudaMemcpy(B.meta_dev, B.meta_host, sizeof(MppMapMeta), cudaMemcpyHostToDevice));
    MPP_CUDA(cudaDeviceSynchronize());
    MPP_CUDA(cudaFree(gdev));
    *out = m;
    return MPP_OK;
}

static uint32_t host_orient_mask(int dR, int dC) {                // MAACO.py:146-157
    uint32_t k = 0xffu;
    if (dC > 0) k &= ~0x29u;
    if (dC < 0) k &= ~0x94u;
    if (dR > 0) k &= ~0x07u;
    if (dR < 0) k &= ~0xE0u;
    return k;
}

MppMapMeta mpp_make_meta(int rows, int cols, int start, int target) {
    (void)rows;
    MppMapMeta M;
    memset(&M, 0, sizeof(M));
    M.start = start; M.target = target;
    if (start < 0 || target < 0) return M;
    MppS1 &S = M.s1;
    S.P1 = host_orient_mask(target / cols - start / cols, target % cols - start % cols);
    S.fast_ok = __builtin_popcount(S.P1) == 3;
    uint32_t rest = S.P1;
    for (int i = 0; i < 3; ++i) {
        const int m = S.fast_ok ? __builtin_ctz(rest) : __builtin_ctz(S.P1 ? S.P1 : 1u);
        rest &= rest - 1;
        const int dr = (int)((0xA940u >> (2 * m)) & 3u) - 1, dc = (int)((0x9224u >> (2 * m)) & 3u) - 1;
        S.sm[i] = m; S.dpr[i] = dr * 8; S.dc[i] = dc; S.so[i] = (dr * cols + dc) * 9 + m + 1;
    }
    return M;
}

extern "C" const mpp_map_batch *mpp_map_as_batch(const mpp_map *m) { return m ? &m->self_batch : nullptr; }

// A batch of n_maps same-shape maps (BASELINE config 5: independent maps, one launch per colony pass of the
// whole batch).  grids_host: n_maps * rows * cols bytes.
extern "C" int mpp_map_batch_create(const uint8_t *grids_host, int n_maps, int rows, int cols, int device,
                                    mpp_map_batch **out) {
    MPP_REQUIRE(grids_host && out, "mpp_map_batch_create: null argument");
    MPP_REQUIRE(n_maps > 0 && n_maps <= 65535, "mpp_map_batch_create: n_maps=%d (1..65535)", n_maps);
    MPP_REQUIRE(rows > 0 && cols > 0 && (long long)rows * cols < (1ll << 30), "mpp_map_batch_create: bad shape %dx%d", rows, cols);
    int rc = mpp_check_device(device);
    if (rc) return rc;
    MPP_CUDA(cudaSetDevice(device));
    mpp_map_batch *B = (mpp_map_batch *)calloc(1, sizeof(mpp_map_batch));
    if (!B) { mpp_set_error("out of host memory"); return MPP_ENOMEM; }
    B->n_maps = n_maps; B->rows = rows; B->cols = cols; B->device = device; B->owns = 1;
    const size_t n = (size_t)rows * cols;
    B->grid_host = (uint8_t *)malloc(n * n_maps);
    B->meta_host = (MppMapMeta *)malloc(sizeof(MppMapMeta) * n_maps);
    if (!B->grid_host || !B->meta_host) { mpp_set_error("out of host memory"); return MPP_ENOMEM; }
    memcpy(B->grid_host, grids_host, n * n_maps);
    for (int k = 0; k < n_maps; ++k) {
        const uint8_t *g = grids_host + n * k;
        int start = -1, target = -1;
        for (size_t i = 0; i < n; ++i) {
            if (g[i] == 2 && start < 0) start = (int)i;             // first row-major hit, MAACO.py:32-41
            if (g[i] == 3 && target < 0) target = (int)i;
        }
        B->meta_host[k] = mpp_make_meta(rows, cols, start, target);
    }
    B->pitch_words = (cols + 2 + 31) / 32;
    B->occ_words = ((rows + 2) * B->pitch_words + 3) & ~3;
    MPP_CUDA(cudaDeviceGetAttribute(&B->sm_count, cudaDevAttrMultiProcessorCount, device));
    uint8_t *gdev = nullptr;
    MPP_CUDA(cudaMalloc(&gdev, n * n_maps));
    MPP_CUDA(cudaMalloc(&B->occ_dev, (size_t)B->occ_words * 4 * n_maps));
    MPP_CUDA(cudaMalloc(&B->svalid_dev, n * n_maps));
    MPP_CUDA(cudaMalloc(&B->meta_dev, sizeof(MppMapMeta) * n_maps));
    MPP_CUDA(cudaMemcpy(gdev, grids_host, n * n_maps, cudaMemcpyHostToDevice));
    MPP_CUDA(cudaMemcpy(B->meta_dev, B->meta_host, sizeof(MppMapMeta) * n_maps, cudaMemcpyHostToDevice));
    mpp_pack_occ_kernel<<<dim3((B->occ_words + 255) / 256, n_maps), 256>>>(gdev, rows, cols, B->pitch_words, B->occ_words,
                                                                           B->occ_dev);
    MPP_CUDA(cudaGetLastError());
    mpp_svalid_kernel<<<dim3(((int)n + 255) / 256, n_maps), 256>>>(B->occ_dev, B->occ_words, B->pitch_words, rows, cols,
                                                                   B->svalid_dev);
    MPP_CUDA(cudaGetLastError());
    MPP_CUDA(cudaDeviceSynchronize());
    MPP_CUDA(cudaFree(gdev));
    *out = B;
    return MPP_OK;
}

extern "C" void mpp_map_batch_destroy(mpp_map_batch *B) {
    if (!B || !B->owns) return;
    cudaSetDevice(B->device);
    if (B->occ_dev) cudaFree(B->occ_dev);
    if (B->svalid_dev) cudaFree(B->svalid_dev);
    if (B->meta_dev) cudaFree(B->meta_dev);
    free(B->grid_host);
    free(B->meta_host);
    free(B);
}
extern "C" int mpp_map_batch_size(const mpp_map_batch *B) { return B ? B->n_maps : -1; }
extern "C" int mpp_map_batch_start(const mpp_map_batch *B, int k) { return (B && k >= 0 && k < B->n_maps) ? B->meta_host[k].start : -1; }
extern "C" int mpp_map_batch_target(const mpp_map_batch *B, int k) { return (B && k >= 0 && k < B->n_maps) ? B->meta_host[k].target : -1; }

extern "C" void mpp_map_destroy(mpp_map *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->occ_dev) cudaFree(m->occ_dev);
    if (m->svalid_dev) cudaFree(m->svalid_dev);
    if (m->safety_d2_dev) cudaFree(m->safety_d2_dev);
    if (m->safety_lut_dev) cudaFree(m->safety_lut_dev);
    if (m->self_batch.meta_dev) cudaFree(m->self_batch.meta_dev);
    free(m->self_batch.meta_host);
    free(m->grid_host);
    free(m);
}

extern "C" int mpp_map_rows(const mpp_map *m) { return m ? m->rows : -1; }
extern "C" int mpp_map_cols(const mpp_map *m) { return m ? m->cols : -1; }
extern "C" int mpp_map_start(const mpp_map *m) { return m ? m->start : -1; }
extern "C" int mpp_map_target(const mpp_map *m) { return m ? m->target : -1; }
extern "C" int mpp_map_device(const mpp_map *m) { return m ? m->device : -1; }
extern "C" const uint32_t *mpp_map_occ_bits(const mpp_map *m, int *pitch_words) {
    if (!m) return nullptr;
    if (pitch_words) *pitch_words = m->pitch_words;
    return m->occ_dev;
}
