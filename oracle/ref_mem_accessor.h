// TEST INFRASTRUCTURE ONLY.  In-memory subclass of the reference's abstract DataAccessor
// (components/iscesys/ImageApi/DataAccessor/include/DataAccessor.h:55-120): lets the reference's own C++ sources, compiled
// unchanged where they lie (oracle/Makefile, target ref), read and write plain buffers instead of image files.
// Lines are pixel-interleaved, as every accessor of the reference's ImageApi delivers them.
#ifndef REF_MEM_ACCESSOR_H
#define REF_MEM_ACCESSOR_H
#include <cstring>

#include "DataAccessor.h"

class MemAccessor : public DataAccessor {
  public:
    MemAccessor(void *base, int lines, int width, int bands, int elsize)
        : base_((char *)base), line_bytes_((size_t)width * bands * elsize), rd_(0), wr_(0)
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
    int getLineSequential(char *buf) override
    {
        if (rd_ >= NumberOfLines) return -1;
        memcpy(buf, base_ + (size_t)rd_ * line_bytes_, line_bytes_);
        return rd_++;
    }
    void setLineSequential(char *buf) override
    {
        if (wr_ < NumberOfLines) memcpy(base_ + (size_t)wr_ * line_bytes_, buf, line_bytes_);
        wr_++;
    }
    double getPx2d(int, int) override { return 0.0; }
    double getPx1d(int) override { return 0.0; }
    int getLineBand(char *, int, int) override { return -1; }
    void setLineBand(char *, int, int) override {}
    void setLineSequentialBand(char *, int) override {}
    void setStream(char *, int &) override {}
    void setStreamAtPos(char *, int &, int &) override {}
    void setSequentialElements(char *, int, int, int) override {}
    void getStream(char *, int &) override {}
    void getStreamAtPos(char *, int &, int &) override {}
    void getSequentialElements(char *, int, int, int &) override {}
    int getLineSequentialBand(char *, int) override { return -1; }
    void finalize() override {}

  private:
    char *base_;
    size_t line_bytes_;
    int rd_, wr_;
};
#endif
