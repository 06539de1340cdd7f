// Real spherical harmonics of the (unit) view direction, degree 1..8, forward (+ dy_dx) and input-gradient.
//
// Behavioural contract: the reference's shencoder extension (ref = /root/reference/im2scene/sdf/models/shencoder):
//   forward  + dy_dx   ref src/shencoder.cu:27-355, launcher :384-397, entry :400-416
//   backward           ref src/shencoder.cu:358-382, entry :419-438
//   layouts            outputs [N, C*C]; dy_dx [N, 3, C*C] (d/dx block, d/dy block, d/dz block)  ref :37-41,126-128
//
// The reference spells out 64 generated polynomials.  All of them are   Y[l*l+l+-m] = A(l,m) * Q_lm(z) * {Re,Im}(x+iy)^m
// with Q_lm = d^m/dz^m P_l(z) (a polynomial in z of degree l-m) and A(l,m) = (-1)^m K(l,m); the kernel evaluates that
// definition with a host-built coefficient table (Horner in z, complex powers by recurrence) instead of 64 unrolled
// expressions, so one code path serves every degree.  In the renderer the view direction is identical for all samples
// of a ray (ref sdf_model.py:304), so callers encode once per RAY and the field kernels broadcast it -- the reference
// evaluates it 24x redundantly per ray.
#include <cmath>

#include "common.cuh"

namespace sdfg {

constexpr int kMaxDeg = 8;
constexpr int kPairs = kMaxDeg * (kMaxDeg + 1) / 2;   // (l,m) pairs with 0 <= m <= l < 8

struct ShTable {
    float q[kPairs][kMaxDeg];   // A(l,m) * coefficients of d^m/dz^m P_l, ascending powers of z
};

static ShTable build_table() {
    double P[kMaxDeg][kMaxDeg] = {};
    P[0][0] = 1.0;
    P[1][1] = 1.0;
    for (int l = 1; l + 1 < kMaxDeg; l++)           // Bonnet: (l+1) P_{l+1} = (2l+1) z P_l - l P_{l-1}
        for (int k = 0; k < kMaxDeg; k++) {
            const double a = k > 0 ? (2 * l + 1) * P[l][k - 1] : 0.0;
            P[l + 1][k] = (a - l * P[l - 1][k]) / (l + 1);
        }
    auto fact = [](int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; };
    ShTable t;
    const double pi = 3.14159265358979323846;
    for (int l = 0; l < kMaxDeg; l++) {
        double q[kMaxDeg];
        for (int k = 0; k < kMaxDeg; k++) q[k] = P[l][k];
        for (int m = 0; m <= l; m++) {
            const double K = std::sqrt((2 * l + 1) / (4 * pi) * fact(l - m) / fact(l + m)) * (m ? std::sqrt(2.0) : 1.0);
            const double a = (m & 1) ? -K : K;
            for (int k = 0; k < kMaxDeg; k++) t.q[l * (l + 1) / 2 + m][k] = (float)(a * q[k]);
            for (int k = 0; k + 1 < kMaxDeg; k++) q[k] = (k + 1) * q[k + 1];   // differentiate for the next m
            q[kMaxDeg - 1] = 0;
        }
    }
    return t;
}

static const ShTable& table() {
    static const ShTable t = build_table();
    return t;
}

template <bool DYDX>
__global__ void __launch_bounds__(256) sh_forward_kernel(const float* __restrict__ inputs, float* __restrict__ outputs,
                                                         float* __restrict__ dy_dx, uint32_t N, int C,
                                                         const __grid_constant__ ShTable tab) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float x = __ldg(inputs + n * 3 + 0), y = __ldg(inputs + n * 3 + 1), z = __ldg(inputs + n * 3 + 2);
    const int C2 = C * C;
    float* o = outputs + n * C2;
    float* dx = DYDX ? dy_dx + n * 3 * C2 : nullptr;
    float* dy = DYDX ? dx + C2 : nullptr;
    float* dz = DYDX ? dy + C2 : nullptr;
    float re[kMaxDeg], im[kMaxDeg];   // (x+iy)^m
    re[0] = 1.f; im[0] = 0.f;
#pragma unroll
    for (int m = 1; m < kMaxDeg; m++) { re[m] = re[m - 1] * x - im[m - 1] * y; im[m] = re[m - 1] * y + im[m - 1] * x; }
    for (int l = 0; l < C; l++) {
        for (int m = 0; m <= l; m++) {
            const float* q = tab.q[l * (l + 1) / 2 + m];
            float Q = 0.f, dQ = 0.f;
#pragma unroll
            for (int k = kMaxDeg - 1; k >= 0; k--) {
                if (DYDX) dQ = fmaf(dQ, z, Q);    // derivative by Horner on the running value
                Q = fmaf(Q, z, q[k]);
            }
            const int ip = l * l + l + m, in = l * l + l - m;
            if (m == 0) {
                o[ip] = Q;
                if (DYDX) { dx[ip] = 0.f; dy[ip] = 0.f; dz[ip] = dQ; }
            } else {
                o[ip] = Q * re[m];
                o[in] = Q * im[m];
                if (DYDX) {
                    const float mq = (float)m * Q;
                    dx[ip] = mq * re[m - 1];  dy[ip] = -mq * im[m - 1];  dz[ip] = dQ * re[m];
                    dx[in] = mq * im[m - 1];  dy[in] = mq * re[m - 1];   dz[in] = dQ * im[m];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) sh_backward_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                                          float* __restrict__ grad_inputs, uint32_t N, int C2) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)N * 3) return;
    const size_t n = t / 3;
    const float* g = grad + n * C2;
    const float* d = dy_dx + t * C2;
    float acc = 0.f;
    for (int ch = 0; ch < C2; ch++) acc = fmaf(__ldg(g + ch), __ldg(d + ch), acc);
    grad_inputs[t] += acc;
}

}  // namespace sdfg

using namespace sdfg;

extern "C" int sdfg_sh_encode_forward(const float* inputs, float* outputs, uint32_t N, uint32_t degree, float* dy_dx,
                                      void* stream) {
    SDFG_REQUIRE(degree >= 1 && degree <= (uint32_t)kMaxDeg, SDFG_ERR_UNSUPPORTED, "sh_encode: degree must be in 1..8 (got %u)", degree);
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(inputs && outputs, SDFG_ERR_INVALID, "sh_encode_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = ceil_div<uint32_t>(N, 256);
    if (dy_dx) sh_forward_kernel<true><<<blocks, 256, 0, st>>>(inputs, outputs, dy_dx, N, (int)degree, table());
    else sh_forward_kernel<false><<<blocks, 256, 0, st>>>(inputs, outputs, nullptr, N, (int)degree, table());
    return check_launch("sh_forward_kernel");
}

extern "C" int sdfg_sh_encode_backward(const float* grad, const float* dy_dx, float* grad_inputs, uint32_t N, uint32_t degree,
                                       void* stream) {
    SDFG_REQUIRE(degree >= 1 && degree <= (uint32_t)kMaxDeg, SDFG_ERR_UNSUPPORTED, "sh_encode: degree must be in 1..8 (got %u)", degree);
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(grad && dy_dx && grad_inputs, SDFG_ERR_INVALID, "sh_encode_backward: null pointer");
    sh_backward_kernel<<<(unsigned)ceil_div<size_t>((size_t)N * 3, 256), 256, 0, (cudaStream_t)stream>>>(
        grad, dy_dx, grad_inputs, N, (int)(degree * degree));
    return check_launch("sh_backward_kernel");
}
