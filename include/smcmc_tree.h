// smcmc_tree.h -- the TTree the sampler writes to.
//
// With ROOT available (<TTree.h> on the include path) this header is just
// `#include <TTree.h>`.  Without ROOT it provides a minimal in-memory TTree
// with the calls the sampler header makes (TSimpleMCMC.H:208-216, :300-349,
// :1517-1596, :1616-1626): Branch() binds the address of a double, an int or
// a std::vector<double>; Fill() snapshots every bound object.
#ifndef SMCMC_TREE_H_SEEN
#define SMCMC_TREE_H_SEEN

#if defined(__has_include)
#if __has_include(<TTree.h>)
#define SMCMC_HAVE_ROOT_TTREE 1
#endif
#endif

#ifdef SMCMC_HAVE_ROOT_TTREE
#include <TTree.h>
namespace sMCMC { namespace detail {
inline bool TreeHasBranch(TTree* t, const char* name) { return t->GetBranch(name) != NULL; }
} }
#else
#include <map>
#include <string>
#include <vector>

#ifndef NULL
#define NULL 0
#endif

class TTree {
public:
    TTree(const char* name = "", const char* title = "") : fName(name), fTitle(title), fEntries(0) {}
    const char* GetName() const { return fName.c_str(); }
    void Branch(const char* n, double* a) { fD[n].src = a; }
    void Branch(const char* n, int* a) { fI[n].src = a; }
    void Branch(const char* n, std::vector<double>* a) { fV[n].src = a; }
    int Fill() {
        for (auto& c : fD) c.second.rows.push_back(*c.second.src);
        for (auto& c : fI) c.second.rows.push_back(*c.second.src);
        for (auto& c : fV) c.second.rows.push_back(*c.second.src);
        return (int)++fEntries;
    }
    long GetEntries() const { return fEntries; }
    int Write(const char* = 0) { return 0; }
    // Reading the way the reference does (TSimpleMCMC.H:300-349, :1517-1596):
    // SetBranchAddress binds a destination (for vectors: the address of a
    // pointer to the vector, as ROOT wants it), GetEntry(i) fills every bound
    // destination, SetBranchAddress(name, NULL) detaches.
    template <class T>
    void SetBranchAddress(const char* n, T* a) { BindRead(n, a); }
    void SetBranchAddress(const char* n, void*) {
        if (fD.count(n)) fD[n].dst = 0;
        if (fI.count(n)) fI[n].dst = 0;
        if (fV.count(n)) fV[n].dst = 0;
    }
    int GetEntry(long i) {
        if (i < 0 || i >= fEntries) return 0;
        for (auto& c : fD) if (c.second.dst) *c.second.dst = c.second.rows[i];
        for (auto& c : fI) if (c.second.dst) *c.second.dst = c.second.rows[i];
        for (auto& c : fV) if (c.second.dst && *c.second.dst) **c.second.dst = c.second.rows[i];
        return 1;
    }
    // Reading back (tests, Restore): column access by name.
    const std::vector<double>& DoubleColumn(const std::string& n) const { return fD.at(n).rows; }
    const std::vector<int>& IntColumn(const std::string& n) const { return fI.at(n).rows; }
    const std::vector<std::vector<double> >& VectorColumn(const std::string& n) const { return fV.at(n).rows; }
    bool HasBranch(const std::string& n) const { return fD.count(n) || fI.count(n) || fV.count(n); }

private:
    template <class T, class Dst = T*>
    struct Column {
        const T* src = 0;
        Dst dst = 0;
        std::vector<T> rows;
    };
    void BindRead(const char* n, double* a) { if (fD.count(n)) fD[n].dst = a; }
    void BindRead(const char* n, int* a) { if (fI.count(n)) fI[n].dst = a; }
    void BindRead(const char* n, std::vector<double>** a) { if (fV.count(n)) fV[n].dst = a; }
    std::string fName, fTitle;
    long fEntries;
    std::map<std::string, Column<double> > fD;
    std::map<std::string, Column<int> > fI;
    std::map<std::string, Column<std::vector<double>, std::vector<double>**> > fV;
};
namespace sMCMC { namespace detail {
inline bool TreeHasBranch(TTree* t, const char* name) { return t->HasBranch(name); }
} }
#endif
#endif
