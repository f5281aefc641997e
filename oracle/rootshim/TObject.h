// oracle/rootshim -- a minimal stand-in for the parts of CERN ROOT that the
// reference headers touch (SURVEY.md section 8c lists the surface).  ROOT is
// not installed here; these headers let /root/reference/*.H compile
// UNMODIFIED so that the reference's own arithmetic is the parity oracle.
// TEST INFRASTRUCTURE ONLY: nothing in the product links or includes this.
#ifndef ROOTSHIM_TObject_h
#define ROOTSHIM_TObject_h
#include <string>
class TObject {
public:
    virtual ~TObject() {}
    virtual int Write(const char* = 0) { return 0; }
};
enum EColor { kWhite = 0, kBlack = 1, kGray = 920, kRed = 632, kGreen = 416,
              kBlue = 600, kYellow = 400, kMagenta = 616, kCyan = 432 };
#ifndef NULL
#define NULL 0
#endif
#endif
