// Poisson upper tail p = 1 - pdtr(k, mu) in fp64, the quantity the reference computes with
// scipy.stats.poisson(mu).cdf(k) (/root/reference/hicpeaks/callers.py:268-270 and :536-540).
//
// scipy evaluates pdtr(k, mu) as the regularised upper incomplete gamma function Q(k+1, mu) (cephes
// igamc, not shipped in /root/reference).  The published algorithm is: for x < a the lower series
// P(a, x) = x^a e^-x / Gamma(a+1) * sum_n x^n / ((a+1)...(a+n)) and Q = 1 - P; otherwise the
// continued fraction for Q; the prefactor x^a e^-x / Gamma(a) is evaluated in a cancellation-free
// form.  The final p = 1 - Q keeps the reference's double subtraction on purpose: for significant
// pixels (k >> mu) the reference's p is quantised to multiples of 2^-53 by "1 - (1 - P)" and the
// same quantisation is needed to agree with it.  Here a = k + 1 is always a positive integer.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define HP_HD __host__ __device__
#else
#define HP_HD
#endif

namespace hp {

HP_HD inline double log1pmx(double u) {  // log(1 + u) - u
    if (fabs(u) < 0.5) {
        double xfac = u, res = 0.0;
        for (int n = 2; n < 600; ++n) {
            xfac *= -u;
            double t = xfac / (double)n;
            res += t;
            if (fabs(t) < 1.1102230246251565e-16 * fabs(res)) break;
        }
        return res;
    }
    return log1p(u) - u;
}

// x^a e^-x / Gamma(a), a integer-valued >= 1
HP_HD inline double igam_fac(double a, double x) {
    if (a < 10.5) {
        double f = 1.0;  // (a-1)!
        for (int j = 2; j < (int)(a + 0.5); ++j) f *= (double)j;
        double t = a * log(x) - x;
        if (t < -745.0) return 0.0;
        return exp(t) / f;
    }
    // Gamma(a) = sqrt(2 pi / a) (a/e)^a exp(c(a)),  c = Stirling series
    double u = (x - a) / a;
    double ia = 1.0 / a, ia2 = ia * ia;
    double c = ia * (1.0 / 12 + ia2 * (-1.0 / 360 + ia2 * (1.0 / 1260 + ia2 * (-1.0 / 1680 + ia2 * (1.0 / 1188 +
               ia2 * (-691.0 / 360360 + ia2 * (1.0 / 156 + ia2 * (-3617.0 / 122400))))))));
    double t = a * log1pmx(u) - c;
    if (t < -745.0) return 0.0;
    return sqrt(a * 0.15915494309189535) * exp(t);  // sqrt(a / (2 pi))
}

// p = 1 - Q(a, x),  a = floor(k) + 1
HP_HD inline double poisson_sf(double k, double mu) {
    if (!(mu > 0.0)) return (mu == 0.0 && k >= 0.0) ? 0.0 : nan("");
    if (k < 0.0) return 1.0;
    double a = floor(k) + 1.0, x = mu;
    double fac = igam_fac(a, x);
    double Q;
    if (x < a) {
        double P = 0.0;
        if (fac != 0.0) {
            double r = a, c = 1.0, ans = 1.0;
            for (int n = 0; n < 2000000; ++n) {
                r += 1.0;
                c *= x / r;
                ans += c;
                if (c <= 1.1102230246251565e-16 * ans) break;
            }
            P = ans * fac / a;
        }
        Q = 1.0 - P;
    } else {
        if (fac == 0.0) {
            Q = 0.0;
        } else if (a == 1.0) {
            Q = exp(-x);
        } else {
            const double big = 4503599627370496.0, biginv = 2.22044604925031308085e-16;
            double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
            double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
            double ans = pkm1 / qkm1, t;
            int it = 0;
            do {
                c += 1.0; y += 1.0; z += 2.0;
                double yc = y * c;
                double pk = pkm1 * z - pkm2 * yc;
                double qk = qkm1 * z - qkm2 * yc;
                if (qk != 0.0) {
                    double r = pk / qk;
                    t = fabs((ans - r) / r);
                    ans = r;
                } else {
                    t = 1.0;
                }
                pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
                if (fabs(pk) > big) { pkm2 *= biginv; pkm1 *= biginv; qkm2 *= biginv; qkm1 *= biginv; }
                if (yc == 0.0) break;  // integer a: the fraction terminates
            } while (t > 1.1102230246251565e-16 && ++it < 2000000);
            Q = ans * fac;
        }
    }
    return 1.0 - Q;
}

}  // namespace hp
