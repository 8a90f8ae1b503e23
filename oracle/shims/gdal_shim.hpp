// TEST INFRASTRUCTURE ONLY -- a stand-in for the part of GDAL the reference's block drivers call, so that
// src/nmap/nmap.cpp, src/evd/evd.cpp, src/phase_link/phase_link.cpp, src/despeck/despeck.cpp,
// src/ampdispersion/ampdispersion.cpp and src/calamp/calamp.cpp compile UNMODIFIED where GDAL is not installed
// (oracle/ref_drivers/*.cpp include them textually, the way the reference's own Cython modules do).
// Files are read and written through the repository's raster layer (fringe_b200/csrc/host/raster_io.hpp: the VRT / ENVI
// subset FRInGE itself produces) -- pure I/O, no arithmetic of the path.  What GDAL does to the numbers on this path is
// sample-type conversion in RasterIO (e.g. the Int32 neighbour counts written to an Int16 raster: clamped); that is
// reproduced in convert_samples().
#pragma once
#include <sys/stat.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "raster_io.hpp"

typedef enum { GDT_Unknown = 0, GDT_Byte = 1, GDT_UInt16 = 2, GDT_Int16 = 3, GDT_UInt32 = 4, GDT_Int32 = 5, GDT_Float32 = 6,
               GDT_Float64 = 7, GDT_CInt16 = 8, GDT_CInt32 = 9, GDT_CFloat32 = 10, GDT_CFloat64 = 11 } GDALDataType;
typedef enum { GA_ReadOnly = 0, GA_Update = 1 } GDALAccess;
typedef enum { GF_Read = 0, GF_Write = 1 } GDALRWFlag;
typedef enum { CE_None = 0, CE_Debug = 1, CE_Warning = 2, CE_Failure = 3, CE_Fatal = 4 } CPLErr;
typedef long long GSpacing;
typedef void* GDALDatasetH;
typedef void* GDALRasterBandH;
typedef void* GDALDriverH;
typedef void* GDALMajorObjectH;
struct GDALRasterIOExtraArg;
typedef int (*GDALProgressFunc)(double, const char*, void*);

namespace gdal_shim {

inline int type_bytes(GDALDataType t) {
    switch (t) {
        case GDT_Byte: return 1;
        case GDT_UInt16: case GDT_Int16: return 2;
        case GDT_UInt32: case GDT_Int32: case GDT_Float32: case GDT_CInt16: return 4;
        case GDT_Float64: case GDT_CInt32: case GDT_CFloat32: return 8;
        case GDT_CFloat64: return 16;
        default: return 0;
    }
}
inline bool is_complex(GDALDataType t) { return t >= GDT_CInt16; }
inline GDALDataType from_envi(int t) {
    switch (t) { case 1: return GDT_Byte; case 2: return GDT_Int16; case 3: return GDT_Int32; case 4: return GDT_Float32;
                 case 5: return GDT_Float64; case 6: return GDT_CFloat32; case 9: return GDT_CFloat64; case 12: return GDT_UInt16;
                 case 13: return GDT_UInt32; }
    return GDT_Unknown;
}
inline int to_envi(GDALDataType t) {
    switch (t) { case GDT_Byte: return 1; case GDT_Int16: return 2; case GDT_Int32: return 3; case GDT_Float32: return 4;
                 case GDT_Float64: return 5; case GDT_CFloat32: return 6; case GDT_CFloat64: return 9; case GDT_UInt16: return 12;
                 case GDT_UInt32: return 13; default: return 0; }
}
inline GDALDataType from_name(const std::string& n) {
    const std::string l = fringe_host::lower(n);
    if (l == "byte") return GDT_Byte; if (l == "int16") return GDT_Int16; if (l == "uint16") return GDT_UInt16;
    if (l == "int32") return GDT_Int32; if (l == "uint32") return GDT_UInt32; if (l == "float32") return GDT_Float32;
    if (l == "float64") return GDT_Float64; if (l == "cfloat32") return GDT_CFloat32; if (l == "cfloat64") return GDT_CFloat64;
    return GDT_Unknown;
}

// one sample as (re, im) in double and back, the way GDALCopyWords converts: integers are rounded to nearest and clamped
// to the destination's range, float -> float is a cast, real <- complex keeps the real part
inline void load_sample(const char* p, GDALDataType t, double& re, double& im) {
    im = 0.0;
    switch (t) {
        case GDT_Byte: re = *reinterpret_cast<const uint8_t*>(p); break;
        case GDT_UInt16: { uint16_t v; std::memcpy(&v, p, 2); re = v; break; }
        case GDT_Int16: { int16_t v; std::memcpy(&v, p, 2); re = v; break; }
        case GDT_UInt32: { uint32_t v; std::memcpy(&v, p, 4); re = v; break; }
        case GDT_Int32: { int32_t v; std::memcpy(&v, p, 4); re = v; break; }
        case GDT_Float32: { float v; std::memcpy(&v, p, 4); re = v; break; }
        case GDT_Float64: { std::memcpy(&re, p, 8); break; }
        case GDT_CFloat32: { float v[2]; std::memcpy(v, p, 8); re = v[0]; im = v[1]; break; }
        case GDT_CFloat64: { double v[2]; std::memcpy(v, p, 16); re = v[0]; im = v[1]; break; }
        default: re = 0.0;
    }
}
template <typename I>
inline I clamp_round(double v) {
    if (std::isnan(v)) return 0;
    const double lo = (double)std::numeric_limits<I>::min(), hi = (double)std::numeric_limits<I>::max();
    const double r = std::floor(v + 0.5);
    return (I)(r < lo ? lo : (r > hi ? hi : r));
}
inline void store_sample(char* p, GDALDataType t, double re, double im) {
    switch (t) {
        case GDT_Byte: { const uint8_t v = clamp_round<uint8_t>(re); *reinterpret_cast<uint8_t*>(p) = v; break; }
        case GDT_UInt16: { const uint16_t v = clamp_round<uint16_t>(re); std::memcpy(p, &v, 2); break; }
        case GDT_Int16: { const int16_t v = clamp_round<int16_t>(re); std::memcpy(p, &v, 2); break; }
        case GDT_UInt32: { const uint32_t v = clamp_round<uint32_t>(re); std::memcpy(p, &v, 4); break; }
        case GDT_Int32: { const int32_t v = clamp_round<int32_t>(re); std::memcpy(p, &v, 4); break; }
        case GDT_Float32: { const float v = (float)re; std::memcpy(p, &v, 4); break; }
        case GDT_Float64: { std::memcpy(p, &re, 8); break; }
        case GDT_CFloat32: { const float v[2] = {(float)re, (float)im}; std::memcpy(p, v, 8); break; }
        case GDT_CFloat64: { const double v[2] = {re, im}; std::memcpy(p, v, 16); break; }
        default: break;
    }
}
// n samples from src (type st, stride ss bytes) to dst (type dt, stride ds bytes)
inline void convert_samples(const char* src, GDALDataType st, long ss, char* dst, GDALDataType dt, long ds, long n) {
    if (st == dt) { const int b = type_bytes(st); for (long i = 0; i < n; ++i) std::memcpy(dst + i * ds, src + i * ss, b); return; }
    for (long i = 0; i < n; ++i) { double re, im; load_sample(src + i * ss, st, re, im); store_sample(dst + i * ds, dt, re, im); }
}

}  // namespace gdal_shim

class GDALDataset;

class GDALMajorObject {
public:
    std::map<std::string, std::map<std::string, std::string> > md;      // domain -> key -> value
    virtual ~GDALMajorObject() {}
    virtual const char* GetMetadataItem(const char* key, const char* domain = "") {
        auto d = md.find(domain ? domain : "");
        if (d == md.end()) return nullptr;
        auto it = d->second.find(key);
        return it == d->second.end() ? nullptr : it->second.c_str();
    }
    virtual CPLErr SetMetadataItem(const char* key, const char* value, const char* domain = "") {
        md[domain ? domain : ""][key] = value ? value : "";
        return CE_None;
    }
};

class GDALRasterBand : public GDALMajorObject {
public:
    GDALDataset* ds = nullptr;
    int index = 0;                                              // 0-based
    CPLErr RasterIO(GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize, void* buf, int bxsize, int bysize,
                    GDALDataType btype, GSpacing pixel_space, GSpacing line_space, GDALRasterIOExtraArg* extra = nullptr);
};

class GDALDataset : public GDALMajorObject {
public:
    fringe_host::Raster in;                                     // read side
    fringe_host::EnviWriter out;                                // write side (ENVI, BIP)
    bool writing = false, memory_only = false;
    int xs = 0, ys = 0, nb = 0;
    GDALDataType ftype = GDT_Unknown;                           // sample type of the file(s)
    std::vector<GDALRasterBand> bands;

    int GetRasterXSize() const { return xs; }
    int GetRasterYSize() const { return ys; }
    int GetRasterCount() const { return nb; }
    GDALRasterBand* GetRasterBand(int b) { return (b >= 1 && b <= nb) ? &bands[b - 1] : nullptr; }
    void make_bands() {
        bands.resize(nb);
        for (int b = 0; b < nb; ++b) { bands[b].ds = this; bands[b].index = b; }
    }
    CPLErr SetMetadataItem(const char* key, const char* value, const char* domain = "") override {
        GDALMajorObject::SetMetadataItem(key, value, domain);
        if (writing && domain && std::string(domain) == "ENVI") out.set_metadata(key, value ? value : "");   // header field
        return CE_None;
    }
    // all bands of a window, caller-defined spacing (the weights raster: pixel-interleaved UInt32 words)
    CPLErr RasterIO(GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize, void* buf, int bxsize, int bysize, GDALDataType btype,
                    int nbands, int* band_map, GSpacing pixel_space, GSpacing line_space, GSpacing band_space,
                    GDALRasterIOExtraArg* extra = nullptr);
};

inline CPLErr GDALRasterBand::RasterIO(GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize, void* buf, int bxsize, int bysize,
                                       GDALDataType btype, GSpacing pixel_space, GSpacing line_space, GDALRasterIOExtraArg*) {
    using namespace gdal_shim;
    GDALDataset& d = *ds;
    if (xoff != 0 || xsize != d.xs || bxsize != xsize || bysize != ysize || yoff < 0 || yoff + ysize > d.ys) return CE_Failure;
    const int bb = type_bytes(btype), fb = type_bytes(d.ftype);
    if (pixel_space == 0) pixel_space = bb;
    if (line_space == 0) line_space = pixel_space * xsize;
    std::vector<char> tmp((size_t)xsize * ysize * fb * (d.writing ? d.nb : 1));
    if (rw == GF_Read) {
        if (d.writing) return CE_Failure;
        if (d.in.interleaved) {                                  // ENVI file: BIP, pick this band
            std::vector<char> all((size_t)xsize * ysize * fb * d.nb);
            if (!d.in.read_interleaved_lines(yoff, ysize, all.data())) return CE_Failure;
            for (int y = 0; y < ysize; ++y)
                convert_samples(all.data() + ((size_t)y * xsize * d.nb + index) * fb, d.ftype, (long)fb * d.nb,
                                (char*)buf + (size_t)y * line_space, btype, (long)pixel_space, xsize);
            return CE_None;
        }
        if (!d.in.read_band_lines(index, yoff, ysize, tmp.data(), fb)) return CE_Failure;
        for (int y = 0; y < ysize; ++y)
            convert_samples(tmp.data() + (size_t)y * xsize * fb, d.ftype, fb, (char*)buf + (size_t)y * line_space, btype,
                            (long)pixel_space, xsize);
        return CE_None;
    }
    if (!d.writing || d.nb != 1) return CE_Failure;              // band-wise writes: single-band outputs only
    for (int y = 0; y < ysize; ++y)
        convert_samples((const char*)buf + (size_t)y * line_space, btype, (long)pixel_space, tmp.data() + (size_t)y * xsize * fb,
                        d.ftype, fb, xsize);
    return d.out.write_lines(yoff, ysize, tmp.data()) ? CE_None : CE_Failure;
}

inline CPLErr GDALDataset::RasterIO(GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize, void* buf, int bxsize, int bysize,
                                    GDALDataType btype, int nbands, int*, GSpacing pixel_space, GSpacing line_space,
                                    GSpacing band_space, GDALRasterIOExtraArg*) {
    using namespace gdal_shim;
    if (xoff != 0 || xsize != xs || bxsize != xsize || bysize != ysize || nbands != nb || yoff < 0 || yoff + ysize > ys) return CE_Failure;
    const int fb = type_bytes(ftype);
    std::vector<char> tmp((size_t)xsize * ysize * fb * nb);     // file order: pixel-interleaved
    if (rw == GF_Read) {
        if (writing || !in.interleaved) return CE_Failure;
        if (!in.read_interleaved_lines(yoff, ysize, tmp.data())) return CE_Failure;
        for (int y = 0; y < ysize; ++y)
            for (int b = 0; b < nb; ++b)
                convert_samples(tmp.data() + ((size_t)y * xsize * nb + b) * fb, ftype, (long)fb * nb,
                                (char*)buf + (size_t)y * line_space + (size_t)b * band_space, btype, (long)pixel_space, xsize);
        return CE_None;
    }
    if (!writing) return CE_Failure;
    for (int y = 0; y < ysize; ++y)
        for (int b = 0; b < nb; ++b)
            convert_samples((const char*)buf + (size_t)y * line_space + (size_t)b * band_space, btype, (long)pixel_space,
                            tmp.data() + ((size_t)y * xsize * nb + b) * fb, ftype, (long)fb * nb, xsize);
    return out.write_lines(yoff, ysize, tmp.data()) ? CE_None : CE_Failure;
}

class GDALDriver : public GDALMajorObject {
public:
    std::string name;
    GDALDataset* Create(const char* fname, int xsize, int ysize, int nbands, GDALDataType t, char** /*options: BIP, SUFFIX=ADD*/) {
        if (name != "ENVI") return nullptr;
        GDALDataset* d = new GDALDataset;
        d->writing = true; d->xs = xsize; d->ys = ysize; d->nb = nbands; d->ftype = t;
        if (!d->out.create(fname, xsize, ysize, nbands, gdal_shim::to_envi(t))) { delete d; return nullptr; }
        d->make_bands();
        return d;
    }
    // calamp.cpp copies the stack VRT and then sets one metadata item per band: the copy lives in memory only, the
    // caller of the shim reads the items back through GetMetadataItem
    GDALDataset* CreateCopy(const char*, GDALDataset* src, int, char**, GDALProgressFunc, void*) {
        GDALDataset* d = new GDALDataset;
        d->memory_only = true; d->xs = src->xs; d->ys = src->ys; d->nb = src->nb; d->ftype = src->ftype;
        d->make_bands();
        for (int b = 0; b < d->nb; ++b) d->bands[b].md = src->bands[b].md;
        gdal_shim_last_copy() = d;
        return d;
    }
    static GDALDataset*& gdal_shim_last_copy() { static GDALDataset* p = nullptr; return p; }
};

inline void GDALAllRegister() {}
inline void GDALDestroyDriverManager() {}
inline GDALDatasetH GDALOpen(const char* path, GDALAccess) {
    GDALDataset* d = new GDALDataset;
    if (!d->in.open(path)) { delete d; return nullptr; }
    d->xs = d->in.cols; d->ys = d->in.rows; d->nb = d->in.count();
    d->make_bands();
    if (d->in.interleaved) {
        d->ftype = gdal_shim::from_envi(d->in.envi.data_type);
        for (auto& kv : d->in.envi.fields) {                    // ENVI header fields = the "ENVI" metadata domain
            std::string key = kv.first;
            for (auto& c : key) c = (char)std::toupper((unsigned char)c);
            d->md["ENVI"][key] = kv.second;
        }
    } else {
        d->ftype = gdal_shim::from_name(d->in.bands[0].dtype);
        for (int b = 0; b < d->nb; ++b)
            for (auto& kv : d->in.bands[b].md_slc) d->bands[b].md["slc"][kv.first] = kv.second;
    }
    return d;
}
inline GDALDatasetH GDALOpenShared(const char* path, GDALAccess a) { return GDALOpen(path, a); }
inline void GDALClose(GDALDatasetH h) {
    GDALDataset* d = static_cast<GDALDataset*>(h);
    if (!d) return;
    if (d->memory_only) return;                                  // kept for the caller (see CreateCopy)
    if (d->writing) d->out.close_file();
    delete d;
}
inline GDALDriverH GDALGetDriverByName(const char* n) {
    static GDALDriver envi, vrt;
    envi.name = "ENVI"; vrt.name = "VRT";
    const std::string s(n ? n : "");
    return s == "ENVI" ? &envi : (s == "VRT" ? &vrt : nullptr);
}
inline CPLErr GDALSetMetadataItem(GDALMajorObjectH h, const char* key, const char* value, const char* domain) {
    return static_cast<GDALMajorObject*>(static_cast<GDALRasterBand*>(h))->SetMetadataItem(key, value, domain);
}
inline int GDALTermProgress(double, const char*, void*) { return 1; }

// ---- CPL / CSL / VSI odds and ends -------------------------------------------------------------
inline char** CSLSetNameValue(char** list, const char*, const char*) { return list; }
inline void CSLDestroy(char**) {}
inline double CPLScanDouble(const char* s, int) { return s ? std::atof(s) : 0.0; }
inline const char* CPLFormFilename(const char* path, const char* basename, const char* ext) {
    static thread_local std::string buf;
    buf = (path && *path) ? std::string(path) + "/" : std::string();
    buf += basename ? basename : "";
    if (ext) buf += ext;
    return buf.c_str();
}
inline char* CPLStrdup(const char* s) { return ::strdup(s); }
typedef struct stat VSIStatBufL;
#define VSI_STAT_EXISTS_FLAG 1
#define VSI_STAT_NATURE_FLAG 2
#define VSI_ISDIR(m) S_ISDIR(m)
inline int VSIStatExL(const char* p, VSIStatBufL* st, int) { return ::stat(p, st); }
inline int VSIMkdir(const char* p, long mode) { return ::mkdir(p, (mode_t)mode); }
