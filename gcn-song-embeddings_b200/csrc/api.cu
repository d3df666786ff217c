// C-ABI plumbing: version, error text, device selection and the ps_gemm dispatcher.
#include "common.cuh"
#include "../../include/pinsage_b200.h"
#include <cstdarg>

char* ps_err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int ps_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ps_err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int ps_gemm_simt_launch(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                        const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                        float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                        const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                        cudaStream_t stream);

int ps_gemm_tc_launch(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                      const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                      float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                      const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                      uint32_t* mask, int64_t ldm, float* p_colsum, cudaStream_t stream);

static int g_gemm_backend = 0;  // 0 = tcgen05 3xTF32 where the shape allows, 1 = CUDA-core fp32 only

extern "C" int ps_gemm_backend(int mode) {
    const int old = g_gemm_backend;
    if (mode == 0 || mode == 1) g_gemm_backend = mode;
    return old;
}

extern "C" int ps_version(void) { return PS_ABI_VERSION; }
extern "C" const char* ps_last_error(void) { return ps_err_buf(); }

extern "C" int ps_set_device(int device) {
    PS_CUDA_CHECK(cudaSetDevice(device));
    return PS_OK;
}

extern "C" int ps_gemm(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                       const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                       float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                       const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                       ps_stream_t stream) {
    return ps_gemm_ex(P, ldp, p_kmajor, p_rows, Q, ldq, q_kmajor, q_rows, C, ldc, M, N, K, bias, act, l2norm, norm_out, accumulate, splits,
                      nullptr, 0, stream);
}

extern "C" int ps_gemm_mask_supported(int64_t M, int64_t N, int64_t K) {
    return g_gemm_backend == 0 && M > 0 && N >= 64 && N % 32 == 0 && K >= 32 && K % 4 == 0 && K <= 8192;
}

extern "C" int ps_gemm_ex(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                          const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                          float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                          const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                          uint32_t* mask, int64_t ld_mask, ps_stream_t stream) {
    if (g_gemm_backend == 0) {
        const int rc = ps_gemm_tc_launch(P, ldp, p_kmajor, p_rows, Q, ldq, q_kmajor, q_rows, C, ldc, M, N, K, bias, act, l2norm,
                                         norm_out, accumulate, splits, mask, ld_mask, nullptr, static_cast<cudaStream_t>(stream));
        if (rc != PS_ERR_UNSUPPORTED) return rc;
    }
    if (mask != nullptr)
        return ps_fail(PS_ERR_UNSUPPORTED, "ps_gemm_ex: the sign-mask epilogue needs the tensor-core path (check ps_gemm_mask_supported)");
    return ps_gemm_simt_launch(P, ldp, p_kmajor, p_rows, Q, ldq, q_kmajor, q_rows, C, ldc, M, N, K, bias, act, l2norm,
                               norm_out, accumulate, splits, static_cast<cudaStream_t>(stream));
}

// Weight gradient of a Linear layer in one call: dW += dY^T X and db += colsum(dY)
// (AddmmBackward of nn.Linear, pinsage_model.py:201,208,259): C[i,j] += sum_r P[r,i] Q[rows(r),j], p_colsum[i] += sum_r P[r,i].
// On the tensor-core path the column sums come from the operand tiles the producers already hold; elsewhere a separate
// column-sum pass runs after the CUDA-core GEMM.
extern "C" int ps_gemm_wgrad(const float* P, int64_t ldp, const float* Q, int64_t ldq, const int32_t* q_rows,
                             float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int splits, float* p_colsum,
                             ps_stream_t stream) {
    if (g_gemm_backend == 0) {
        const int rc = ps_gemm_tc_launch(P, ldp, 0, nullptr, Q, ldq, 0, q_rows, C, ldc, M, N, K, nullptr, 0, 0, nullptr, 1, splits,
                                         nullptr, 0, p_colsum, static_cast<cudaStream_t>(stream));
        if (rc != PS_ERR_UNSUPPORTED) return rc;
    }
    const int rc = ps_gemm_simt_launch(P, ldp, 0, nullptr, Q, ldq, 0, q_rows, C, ldc, M, N, K, nullptr, 0, 0, nullptr, 1, splits,
                                       static_cast<cudaStream_t>(stream));
    if (rc != PS_OK || p_colsum == nullptr) return rc;
    return ps_colsum(P, ldp, K, static_cast<int>(M), p_colsum, stream);
}
