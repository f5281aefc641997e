// rootshim TTree: an in-memory recorder.  Branch() binds the ADDRESS of a
// double, int or std::vector<double>; Fill() snapshots every bound object;
// GetEntry() copies an entry back into whatever SetBranchAddress() bound
// (for vectors ROOT wants the address of a pointer, TSimpleMCMC.H:289-302).
#ifndef ROOTSHIM_TTree_h
#define ROOTSHIM_TTree_h
#include <map>
#include <string>
#include <vector>
#include "TObject.h"
class TFile;
class TTree : public TObject {
public:
    struct Column {
        int kind;                         // 0 double, 1 int, 2 vector<double>
        const void* source;               // bound at Branch()
        void* sink;                       // bound at SetBranchAddress()
        bool sinkIsPtrPtr;
        std::vector<double> d;
        std::vector<int> i;
        std::vector<std::vector<double> > v;
        Column() : kind(0), source(0), sink(0), sinkIsPtrPtr(false) {}
    };
    TTree(const char* name = "", const char* title = "")
        : fName(name), fTitle(title), fEntries(0) {}
    const char* GetName() const { return fName.c_str(); }
    void Branch(const char* n, double* a) { Bind(n, 0, a); }
    void Branch(const char* n, int* a) { Bind(n, 1, a); }
    void Branch(const char* n, std::vector<double>* a) { Bind(n, 2, a); }
    void SetBranchAddress(const char* n, double* a) { Sink(n, a, false); }
    void SetBranchAddress(const char* n, int* a) { Sink(n, a, false); }
    void SetBranchAddress(const char* n, std::vector<double>* a) { Sink(n, a, false); }
    void SetBranchAddress(const char* n, std::vector<double>** a) { Sink(n, a, true); }
    void SetBranchAddress(const char* n, long) { Sink(n, 0, false); }   // NULL
    int Fill() {
        for (std::map<std::string, Column>::iterator c = fColumns.begin();
             c != fColumns.end(); ++c) {
            Column& col = c->second;
            if (col.kind == 0) col.d.push_back(col.source ? *(const double*)col.source : 0.0);
            else if (col.kind == 1) col.i.push_back(col.source ? *(const int*)col.source : 0);
            else col.v.push_back(col.source ? *(const std::vector<double>*)col.source
                                            : std::vector<double>());
        }
        ++fEntries;
        return 1;
    }
    long GetEntries() const { return fEntries; }
    int GetEntry(long e) {
        if (e < 0 || e >= fEntries) return 0;
        for (std::map<std::string, Column>::iterator c = fColumns.begin();
             c != fColumns.end(); ++c) {
            Column& col = c->second;
            if (!col.sink) continue;
            if (col.kind == 0) *(double*)col.sink = col.d[e];
            else if (col.kind == 1) *(int*)col.sink = col.i[e];
            else if (col.sinkIsPtrPtr) **(std::vector<double>**)col.sink = col.v[e];
            else *(std::vector<double>*)col.sink = col.v[e];
        }
        return 1;
    }
    void SetDirectory(TFile*) {}
    const Column* GetColumn(const std::string& n) const {
        std::map<std::string, Column>::const_iterator c = fColumns.find(n);
        return c == fColumns.end() ? 0 : &c->second;
    }
private:
    void Bind(const char* n, int kind, const void* a) {
        Column& col = fColumns[n];
        col.kind = kind; col.source = a;
    }
    void Sink(const char* n, void* a, bool pp) {
        std::map<std::string, Column>::iterator c = fColumns.find(n);
        if (c == fColumns.end()) return;
        c->second.sink = a; c->second.sinkIsPtrPtr = pp;
    }
    std::string fName, fTitle;
    long fEntries;
    std::map<std::string, Column> fColumns;
};
#endif
