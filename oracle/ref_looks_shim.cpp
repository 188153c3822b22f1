// TEST INFRASTRUCTURE ONLY.  Drives the reference's own takeLooks<T> / takeLookscpx<T> templates -- compiled UNCHANGED
// from components/mroipac/looks/bindings/looksmodule.cpp where it lies (see oracle/Makefile, target ref) -- on plain
// memory buffers, through an in-memory subclass of the reference's abstract DataAccessor.  getLine / setLine carry
// pixel-interleaved lines, as every accessor of the reference's ImageApi delivers them (DataAccessorCaster.cpp:5-30,
// BILAccessor.cpp:54-90).  Used by tests/ to pin the numpy restatement oracle.looks().
#include <complex>
#include <cstring>

#include "DataAccessor.h"

template <typename T> int takeLooks(DataAccessor *IAIn, DataAccessor *IAout, int ld, int la);
template <typename T> int takeLookscpx(DataAccessor *IAIn, DataAccessor *IAout, int ld, int la);

namespace {
class MemAccessor : public DataAccessor {
  public:
    MemAccessor(char *base, int lines, int width, int bands, int elsize) : base_(base), line_bytes_((size_t)width * bands * elsize)
    {
        Accessor = nullptr;
        Caster = nullptr;
        DataSizeIn = DataSizeOut = elsize;
        Bands = bands;
        LineWidth = width;
        LineCounter = 0;
        poly = nullptr;
        NumberOfLines = lines;
        LineOffset = 0;
    }
    int getLine(char *buf, int pos) override
    {
        if (pos < 0 || pos >= NumberOfLines) return -1;
        memcpy(buf, base_ + (size_t)pos * line_bytes_, line_bytes_);
        return 0;
    }
    void setLine(char *buf, int pos) override { memcpy(base_ + (size_t)pos * line_bytes_, buf, line_bytes_); }
    double getPx2d(int, int) override { return 0.0; }
    double getPx1d(int) override { return 0.0; }
    int getLineBand(char *, int, int) override { return -1; }
    void setLineBand(char *, int, int) override {}
    void setLineSequential(char *) override {}
    void setLineSequentialBand(char *, int) override {}
    void setStream(char *, int &) override {}
    void setStreamAtPos(char *, int &, int &) override {}
    void setSequentialElements(char *, int, int, int) override {}
    void getStream(char *, int &) override {}
    void getStreamAtPos(char *, int &, int &) override {}
    void getSequentialElements(char *, int, int, int &) override {}
    int getLineSequential(char *) override { return -1; }
    int getLineSequentialBand(char *, int) override { return -1; }
    void finalize() override {}

  private:
    char *base_;
    size_t line_bytes_;
};
} // namespace

// dtype: 0 char, 1 short, 2 int, 3 long, 4 float, 5 double, 6 complex<float>; in: [length][width][bands] (BIP);
// out: [length / ld][width / la][bands]
extern "C" int ref_take_looks(int dtype, const void *in, void *out, int length, int width, int bands, int ld, int la)
{
    static const int sizes[] = {1, 2, 4, 8, 4, 8, 8};
    if (dtype < 0 || dtype > 6) return -1;
    MemAccessor a((char *)in, length, width, bands, sizes[dtype]);
    MemAccessor b((char *)out, length / ld, width / la, bands, sizes[dtype]);
    switch (dtype) {
    case 0: return takeLooks<char>(&a, &b, ld, la);
    case 1: return takeLooks<short>(&a, &b, ld, la);
    case 2: return takeLooks<int>(&a, &b, ld, la);
    case 3: return takeLooks<long>(&a, &b, ld, la);
    case 4: return takeLooks<float>(&a, &b, ld, la);
    case 5: return takeLooks<double>(&a, &b, ld, la);
    default: return takeLookscpx<float>(&a, &b, ld, la);
    }
}
