// rootshim TMatrixD / TMatrixDSym: dense row-major double matrices with the
// few operations the reference calls.  Third-party behaviour restated
// [from memory, SURVEY.md A.6]: Invert() = LU with partial pivoting;
// EigenVectors(values) = symmetric eigen-decomposition, eigenvalues sorted
// DESCENDING, eigenvectors in columns (TSimpleMCMC.H:1287,1301 rely on it).
#ifndef ROOTSHIM_TMatrixD_h
#define ROOTSHIM_TMatrixD_h
#include <algorithm>
#include <cmath>
#include <iostream>
#include <vector>
#include "TVectorD.h"
class TMatrixD {
public:
    TMatrixD() : fR(0), fC(0) {}
    TMatrixD(int r, int c) : fR(r), fC(c), fData((size_t)r * c, 0.0) {}
    void ResizeTo(int r, int c) {
        std::vector<double> nd((size_t)r * c, 0.0);
        for (int i = 0; i < std::min(r, fR); ++i)
            for (int j = 0; j < std::min(c, fC); ++j)
                nd[(size_t)i * c + j] = fData[(size_t)i * fC + j];
        fData.swap(nd); fR = r; fC = c;
    }
    int GetNrows() const { return fR; }
    int GetNcols() const { return fC; }
    double& operator()(int i, int j) { return fData[(size_t)i * fC + j]; }
    double operator()(int i, int j) const { return fData[(size_t)i * fC + j]; }
    const double* GetMatrixArray() const { return fData.data(); }
    double* GetMatrixArray() { return fData.data(); }
    void Print(const char* = "") const {
        for (int i = 0; i < fR; ++i) {
            for (int j = 0; j < fC; ++j) std::cout << " " << (*this)(i, j);
            std::cout << std::endl;
        }
    }
    // In-place inverse, LU with partial pivoting (Doolittle, row swaps).
    TMatrixD& Invert(double* det = 0) {
        const int n = fR;
        std::vector<double> a(fData);
        std::vector<double> inv((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
        double d = 1.0;
        for (int k = 0; k < n; ++k) {
            int piv = k;
            double best = std::abs(a[(size_t)k * n + k]);
            for (int i = k + 1; i < n; ++i) {
                double v = std::abs(a[(size_t)i * n + k]);
                if (v > best) { best = v; piv = i; }
            }
            if (piv != k) {
                for (int j = 0; j < n; ++j) {
                    std::swap(a[(size_t)k * n + j], a[(size_t)piv * n + j]);
                    std::swap(inv[(size_t)k * n + j], inv[(size_t)piv * n + j]);
                }
                d = -d;
            }
            double p = a[(size_t)k * n + k];
            d *= p;
            for (int j = 0; j < n; ++j) {
                a[(size_t)k * n + j] /= p;
                inv[(size_t)k * n + j] /= p;
            }
            for (int i = 0; i < n; ++i) {
                if (i == k) continue;
                double f = a[(size_t)i * n + k];
                if (f == 0.0) continue;
                for (int j = 0; j < n; ++j) {
                    a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
                    inv[(size_t)i * n + j] -= f * inv[(size_t)k * n + j];
                }
            }
        }
        fData.swap(inv);
        if (det) *det = d;
        return *this;
    }
    // Cyclic Jacobi on the symmetric part; descending eigenvalues.
    TMatrixD EigenVectors(TVectorD& values) const {
        const int n = fR;
        std::vector<double> a(fData);
        std::vector<double> v((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
        for (int sweep = 0; sweep < 100; ++sweep) {
            double off = 0.0;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
            if (!(off > 1e-300)) break;
            for (int p = 0; p < n; ++p) {
                for (int q = p + 1; q < n; ++q) {
                    double apq = a[(size_t)p * n + q];
                    if (apq == 0.0) continue;
                    double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                    double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                    double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < n; ++k) {
                        double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                        a[(size_t)k * n + p] = c * akp - s * akq;
                        a[(size_t)k * n + q] = s * akp + c * akq;
                    }
                    for (int k = 0; k < n; ++k) {
                        double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                        a[(size_t)p * n + k] = c * apk - s * aqk;
                        a[(size_t)q * n + k] = s * apk + c * aqk;
                    }
                    for (int k = 0; k < n; ++k) {
                        double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
                        v[(size_t)k * n + p] = c * vkp - s * vkq;
                        v[(size_t)k * n + q] = s * vkp + c * vkq;
                    }
                }
            }
        }
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            return a[(size_t)x * n + x] > a[(size_t)y * n + y];
        });
        values.ResizeTo(n);
        TMatrixD out(n, n);
        for (int c = 0; c < n; ++c) {
            values(c) = a[(size_t)order[c] * n + order[c]];
            for (int r = 0; r < n; ++r) out(r, c) = v[(size_t)r * n + order[c]];
        }
        return out;
    }
protected:
    int fR, fC;
    std::vector<double> fData;
};
class TMatrixDSym : public TMatrixD {
public:
    TMatrixDSym() {}
    explicit TMatrixDSym(int n) : TMatrixD(n, n) {}
};
#endif
