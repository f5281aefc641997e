// rootshim TDecompChol.  ROOT behaviour restated [from memory, SURVEY.md
// A.6]: factor a = U^T U column by column, reading only the upper triangle of
// the input; each pivot is the diagonal minus the running sum of squares of
// the column above it, a pivot <= 0 fails; the rest of the row is the input
// entry minus the running dot product of the two columns above (row index
// ascending), divided by the pivot's square root; the strict lower triangle
// of U is zero.
#ifndef ROOTSHIM_TDecompChol_h
#define ROOTSHIM_TDecompChol_h
#include <cmath>
#include "TMatrixD.h"
class TDecompChol {
public:
    explicit TDecompChol(const TMatrixD& a) : fU(a), fOk(false) {}
    bool Decompose() {
        const int n = fU.GetNrows();
        if (n != fU.GetNcols()) return false;
        for (int c = 0; c < n; ++c) {
            double pivot = fU(c, c);
            for (int r = 0; r < c; ++r) pivot -= fU(r, c) * fU(r, c);
            if (!(pivot > 0.0)) return false;
            pivot = std::sqrt(pivot);
            fU(c, c) = pivot;
            for (int j = c + 1; j < n; ++j) {
                double v = fU(c, j);
                for (int r = 0; r < c; ++r) v -= fU(r, j) * fU(r, c);
                fU(c, j) = v / pivot;
            }
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j) fU(i, j) = 0.0;
        fOk = true;
        return true;
    }
    const TMatrixD& GetU() const { return fU; }
private:
    TMatrixD fU;
    bool fOk;
};
#endif
