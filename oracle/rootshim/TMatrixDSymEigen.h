#ifndef ROOTSHIM_TMatrixDSymEigen_h
#define ROOTSHIM_TMatrixDSymEigen_h
#include "TMatrixD.h"
class TMatrixDSymEigen {
public:
    explicit TMatrixDSymEigen(const TMatrixDSym& m) { fVectors = m.EigenVectors(fValues); }
    const TMatrixD& GetEigenVectors() const { return fVectors; }
    const TVectorD& GetEigenValues() const { return fValues; }
private:
    TMatrixD fVectors;
    TVectorD fValues;
};
#endif
