/*
 * oracle/sh_oracle.c  --  TEST INFRASTRUCTURE ONLY (never linked into / imported by the product).
 *
 * CPU restatement of the reference's real spherical-harmonics view-direction encoder
 * (/root/reference/im2scene/sdf/models/shencoder/src/shencoder.cu:27-355 forward + dy_dx, :358-382 backward).
 *
 * The reference spells out 64 sympy-generated polynomials.  They are all of the canonical form
 *      Y[l*l+l+m] = (-1)^m * K(l,|m|) * Q(l,|m|)(z) * T_m(x,y)
 *   K(l,m) = sqrt((2l+1)/(4 pi) * (l-m)!/(l+m)!) * (m ? sqrt(2) : 1)
 *   Q(l,m) = d^m/dz^m P_l(z)                      (P_l = Legendre polynomial; a polynomial in z only)
 *   T_m    = Re (x+iy)^m  (m>0),  Im (x+iy)^|m|  (m<0),  1 (m=0)
 * (e.g. shencoder.cu:52-54 for l=1, :56-60 for l=2, :62-68 for l=3), and dy_dx is the partial derivative of exactly
 * that representative (shencoder.cu:130-354), NOT of the solid harmonic.  This file evaluates the definition by
 * recurrence in double precision, which is an independent route to the same numbers; it is pinned against the
 * reference's own formulas by tests/golden/sh_deg8.npz (made by tests/golden/make_golden.py, which parses and evaluates
 * the reference's expressions in float32) and, on the GPU box, against oracle/_ref/_shencoder_ref.so.
 *
 * Layouts (shencoder.cu:37-41,126-128): outputs [B, C*C]; dy_dx [B, 3, C*C]  (d/dx block, d/dy block, d/dz block).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define MAXDEG 8

static double factorial(int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; }

/* Legendre polynomial coefficients c[l][k] of z^k via Bonnet: (l+1) P_{l+1} = (2l+1) z P_l - l P_{l-1} */
static void legendre_coeffs(double c[MAXDEG][MAXDEG]) {
    for (int l = 0; l < MAXDEG; l++) for (int k = 0; k < MAXDEG; k++) c[l][k] = 0;
    c[0][0] = 1;
    if (MAXDEG > 1) c[1][1] = 1;
    for (int l = 1; l + 1 < MAXDEG; l++)
        for (int k = 0; k < MAXDEG; k++) {
            double a = (k > 0) ? (2 * l + 1) * c[l][k - 1] : 0.0;
            c[l + 1][k] = (a - l * c[l - 1][k]) / (l + 1);
        }
}

static double poly_eval(const double* c, int n, double z) { double r = 0; for (int k = n - 1; k >= 0; k--) r = r * z + c[k]; return r; }
static void poly_diff(double* c, int n) { for (int k = 0; k + 1 < n; k++) c[k] = (k + 1) * c[k + 1]; c[n - 1] = 0; }

static void sh_point(int C, double x, double y, double z, float* out, float* dx, float* dy, float* dz) {
    static double P[MAXDEG][MAXDEG];
    static int init = 0;
    if (!init) { legendre_coeffs(P); init = 1; }
    /* re[m], im[m] = Re/Im (x+iy)^m */
    double re[MAXDEG + 1], im[MAXDEG + 1];
    re[0] = 1; im[0] = 0;
    for (int m = 1; m <= MAXDEG; m++) { re[m] = re[m - 1] * x - im[m - 1] * y; im[m] = re[m - 1] * y + im[m - 1] * x; }
    for (int l = 0; l < C; l++) {
        double q[MAXDEG];
        for (int k = 0; k < MAXDEG; k++) q[k] = P[l][k];
        for (int m = 0; m <= l; m++) {
            /* q now holds d^m/dz^m P_l */
            double dq[MAXDEG];
            for (int k = 0; k < MAXDEG; k++) dq[k] = q[k];
            poly_diff(dq, MAXDEG);
            const double K = sqrt((2 * l + 1) / (4 * M_PI) * factorial(l - m) / factorial(l + m)) * (m ? sqrt(2.0) : 1.0);
            const double sgn = (m & 1) ? -1.0 : 1.0;
            const double Q = poly_eval(q, MAXDEG, z), dQ = poly_eval(dq, MAXDEG, z);
            const double a = sgn * K;
            const int ip = l * l + l + m, in = l * l + l - m;
            if (m == 0) {
                out[ip] = (float)(a * Q);
                if (dx) { dx[ip] = 0; dy[ip] = 0; dz[ip] = (float)(a * dQ); }
            } else {
                out[ip] = (float)(a * Q * re[m]);
                out[in] = (float)(a * Q * im[m]);
                if (dx) {
                    dx[ip] = (float)(a * Q * m * re[m - 1]);  dy[ip] = (float)(-a * Q * m * im[m - 1]);  dz[ip] = (float)(a * dQ * re[m]);
                    dx[in] = (float)(a * Q * m * im[m - 1]);  dy[in] = (float)(a * Q * m * re[m - 1]);   dz[in] = (float)(a * dQ * im[m]);
                }
            }
            poly_diff(q, MAXDEG);
        }
    }
}

/* shencoder.cu:27-355.  inputs [B,3]; outputs [B,C*C]; dy_dx [B,3,C*C] or NULL.  C = degree in 1..8 */
void oracle_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t C, float* dy_dx) {
    const uint32_t C2 = C * C;
    for (uint32_t b = 0; b < B; b++) {
        const float* p = inputs + (size_t)b * 3;
        float* d = dy_dx ? dy_dx + (size_t)b * 3 * C2 : 0;
        sh_point((int)C, p[0], p[1], p[2], outputs + (size_t)b * C2, d, d ? d + C2 : 0, d ? d + 2 * C2 : 0);
    }
}

/* shencoder.cu:358-382: grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]  (accumulates, caller pre-zeroes) */
void oracle_sh_encode_backward(const float* grad, uint32_t B, uint32_t C, const float* dy_dx, float* grad_inputs) {
    const uint32_t C2 = C * C;
    for (uint32_t b = 0; b < B; b++)
        for (uint32_t d = 0; d < 3; d++) {
            double r = 0;
            for (uint32_t ch = 0; ch < C2; ch++) r += (double)grad[(size_t)b * C2 + ch] * dy_dx[((size_t)b * 3 + d) * C2 + ch];
            grad_inputs[(size_t)b * 3 + d] += (float)r;
        }
}
