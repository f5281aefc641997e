#ifndef ROOTSHIM_TVectorD_h
#define ROOTSHIM_TVectorD_h
#include <vector>
class TVectorD {
public:
    TVectorD() {}
    explicit TVectorD(int n) : fData(n, 0.0) {}
    void ResizeTo(int n) { fData.assign(n, 0.0); }
    int GetNrows() const { return (int)fData.size(); }
    double& operator()(int i) { return fData[i]; }
    double operator()(int i) const { return fData[i]; }
    double& operator[](int i) { return fData[i]; }
    double operator[](int i) const { return fData[i]; }
private:
    std::vector<double> fData;
};
#endif
