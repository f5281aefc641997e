// rootshim TFile: a named bag of objects that lives for the process lifetime.
#ifndef ROOTSHIM_TFile_h
#define ROOTSHIM_TFile_h
#include <map>
#include <string>
#include "TObject.h"
class TFile : public TObject {
public:
    TFile(const char* name = "", const char* = "") : fName(name) {}
    TObject* Get(const char* key) {
        std::map<std::string, TObject*>& bag = Registry()[fName];
        std::map<std::string, TObject*>::iterator it = bag.find(key);
        return it == bag.end() ? 0 : it->second;
    }
    void Put(const char* key, TObject* obj) { Registry()[fName][key] = obj; }
    static std::map<std::string, std::map<std::string, TObject*> >& Registry() {
        static std::map<std::string, std::map<std::string, TObject*> > r;
        return r;
    }
private:
    std::string fName;
};
#endif
