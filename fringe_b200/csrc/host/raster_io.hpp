// Minimal raster I/O for the block drivers: exactly the on-disk subset FRInGE produces and
// consumes on this path, with no GDAL dependency:
//   * stack VRT: one VRTRasterBand per date, SimpleSource -> per-date VRT with a SrcRect crop
//     and <Metadata domain="slc"> (python/tops2vrt.py:201-225, src/sequential/Stack.py:20-34);
//     per-date VRT = one VRTRawRasterBand over a flat CFloat32 file (tops2vrt.py:144-152);
//     a band may also be a VRTRawRasterBand directly, or a SimpleSource over an ENVI file;
//   * ENVI rasters (INTERLEAVE=BIP, SUFFIX=ADD => "<name>.hdr") for every output
//     (src/nmap/nmap.cpp:237-264, src/evd/evd.cpp:266-366) and for the wts / mask inputs.
// With -DFRINGE_WITH_GDAL (fringe_b200/build_host.py adds it when `gdal-config` is on the PATH) a raster the built-in
// reader cannot resolve -- GeoTIFF, non-raw VRT sources, pixel functions -- is opened through GDAL's C API instead and
// read with GDALRasterIO into the same buffers; outputs stay ENVI, as in the reference.  GDAL is not part of the image
// this repository is developed in, so that branch is compiled only where GDAL exists.  Nothing here touches the device.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#ifdef FRINGE_WITH_GDAL
#include <gdal.h>
#include <mutex>
#endif

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace fringe_host {

inline std::string dirname_of(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : (k == 0 ? std::string("/") : p.substr(0, k));
}
inline std::string join_path(const std::string& dir, const std::string& f) {
    if (!f.empty() && f[0] == '/') return f;
    return dir + "/" + f;
}
inline bool file_exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
inline bool slurp(const std::string& path, std::string& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::ostringstream ss; ss << f.rdbuf(); out = ss.str();
    return true;
}

// pread / pwrite until every byte has moved: a single call may transfer less than asked (Linux caps one
// call near 2 GiB; signals interrupt it), which is not an error
inline bool pread_all(int fd, void* dst, size_t bytes, off_t off) {
    char* p = static_cast<char*>(dst);
    while (bytes > 0) {
        const ssize_t k = ::pread(fd, p, bytes > ((size_t)1 << 30) ? ((size_t)1 << 30) : bytes, off);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        if (k == 0) return false;                     // end of file before the block was complete
        p += k; off += k; bytes -= (size_t)k;
    }
    return true;
}
inline bool pwrite_all(int fd, const void* src, size_t bytes, off_t off) {
    const char* p = static_cast<const char*>(src);
    while (bytes > 0) {
        const ssize_t k = ::pwrite(fd, p, bytes > ((size_t)1 << 30) ? ((size_t)1 << 30) : bytes, off);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        p += k; off += k; bytes -= (size_t)k;
    }
    return true;
}

// ---- tiny XML helpers (enough for VRT) ------------------------------------------------------
struct XmlElem { std::string open_tag, body; };   // open_tag = text between '<' and '>'
// find next element <name ...>...</name> (or self-closing) at or after pos; returns end offset or npos
inline size_t xml_next(const std::string& s, const std::string& name, size_t pos, XmlElem& e) {
    while (true) {
        size_t a = s.find("<" + name, pos);
        if (a == std::string::npos) return a;
        const char c = s[a + 1 + name.size()];
        if (c != ' ' && c != '>' && c != '/' && c != '\t' && c != '\n' && c != '\r') { pos = a + 1; continue; }
        size_t b = s.find('>', a);
        if (b == std::string::npos) return b;
        e.open_tag = s.substr(a + 1, b - a - 1);
        if (!e.open_tag.empty() && e.open_tag.back() == '/') { e.body.clear(); return b + 1; }
        const std::string close = "</" + name + ">";
        size_t z = s.find(close, b);
        if (z == std::string::npos) return z;
        e.body = s.substr(b + 1, z - b - 1);
        return z + close.size();
    }
}
inline std::string xml_attr(const std::string& tag, const std::string& key) {
    size_t p = 0;
    while ((p = tag.find(key, p)) != std::string::npos) {
        const bool boundary = (p == 0) || std::isspace((unsigned char)tag[p - 1]);
        size_t q = p + key.size();
        while (q < tag.size() && std::isspace((unsigned char)tag[q])) ++q;
        if (boundary && q < tag.size() && tag[q] == '=') {
            ++q;
            while (q < tag.size() && std::isspace((unsigned char)tag[q])) ++q;
            if (q < tag.size() && (tag[q] == '"' || tag[q] == '\'')) {
                const char qc = tag[q];
                size_t r = tag.find(qc, q + 1);
                if (r != std::string::npos) return tag.substr(q + 1, r - q - 1);
            }
        }
        p += key.size();
    }
    return "";
}
inline bool xml_child_text(const std::string& body, const std::string& name, std::string& text, std::string* tag = nullptr) {
    XmlElem e;
    if (xml_next(body, name, 0, e) == std::string::npos) return false;
    text = trim(e.body);
    if (tag) *tag = e.open_tag;
    return true;
}

// ---- a band backed by a flat file ---------------------------------------------------------------
struct RawBand {
    std::string path;
    long image_offset = 0, pixel_offset = 0, line_offset = 0;
    int src_width = 0, src_height = 0;     // size of the source raster
    int x_off = 0, y_off = 0;              // SrcRect origin inside the source
    int elem_bytes = 8;                    // bytes per sample actually read
    std::string dtype = "CFloat32";
    std::map<std::string, std::string> md_slc;
    int fd = -1;
};

inline int dtype_bytes(const std::string& d) {
    const std::string l = lower(d);
    if (l == "byte") return 1;
    if (l == "int16" || l == "uint16") return 2;
    if (l == "int32" || l == "uint32" || l == "float32") return 4;
    if (l == "cfloat32" || l == "float64") return 8;
    return 0;
}

// ---- ENVI header --------------------------------------------------------------------------------
struct EnviHeader {
    int samples = 0, lines = 0, bands = 0, data_type = 0, byte_order = 0;
    long header_offset = 0;
    std::string interleave = "bsq";
    std::map<std::string, std::string> fields;   // every "key = value" (key lower-cased)
};
inline std::string envi_hdr_path(const std::string& data) {
    if (file_exists(data + ".hdr")) return data + ".hdr";
    const size_t dot = data.find_last_of('.');
    if (dot != std::string::npos && file_exists(data.substr(0, dot) + ".hdr")) return data.substr(0, dot) + ".hdr";
    return "";
}
inline bool read_envi_header(const std::string& hdr, EnviHeader& h) {
    std::string txt;
    if (!slurp(hdr, txt)) return false;
    std::istringstream ss(txt);
    std::string line;
    while (std::getline(ss, line)) {
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        std::string key = lower(trim(line.substr(0, eq))), val = trim(line.substr(eq + 1));
        if (!val.empty() && val[0] == '{') {           // multi-line {...}
            while (val.find('}') == std::string::npos && std::getline(ss, line)) val += " " + trim(line);
            const size_t a = val.find('{'), b = val.rfind('}');
            val = trim(val.substr(a + 1, (b == std::string::npos ? val.size() : b) - a - 1));
        }
        h.fields[key] = val;
    }
    auto geti = [&](const char* k, int def) { auto it = h.fields.find(k); return it == h.fields.end() ? def : std::atoi(it->second.c_str()); };
    h.samples = geti("samples", 0); h.lines = geti("lines", 0); h.bands = geti("bands", 1);
    h.data_type = geti("data type", 0); h.byte_order = geti("byte order", 0);
    h.header_offset = geti("header offset", 0);
    if (h.fields.count("interleave")) h.interleave = lower(h.fields["interleave"]);
    return h.samples > 0 && h.lines > 0;
}
inline int envi_type_bytes(int t) {
    switch (t) { case 1: return 1; case 2: case 12: return 2; case 3: case 4: case 13: return 4; case 5: case 6: return 8; case 9: return 16; }
    return 0;
}

// ---- generic multi-band raster reader (VRT subset or ENVI) ------------------------------------------
struct Raster {
    int cols = 0, rows = 0;
    std::vector<RawBand> bands;                        // VRT-style: one flat file per band
    // ENVI-style interleaved file (wts / mask / single-band products)
    bool interleaved = false;
    std::string path;
    EnviHeader envi;
    int fd = -1;
    std::string error;
#ifdef FRINGE_WITH_GDAL
    GDALDatasetH gdal_ds = nullptr;                    // set when the raster is read through GDAL
    std::mutex gdal_mu;                                // a GDAL dataset handle is not thread-safe; block workers share it
#endif

    ~Raster() { close_all(); }
    void close_all() {
        for (auto& b : bands) if (b.fd >= 0) { ::close(b.fd); b.fd = -1; }
        if (fd >= 0) { ::close(fd); fd = -1; }
#ifdef FRINGE_WITH_GDAL
        if (gdal_ds) { GDALClose(gdal_ds); gdal_ds = nullptr; }
#endif
    }
    int count() const { return interleaved ? envi.bands : (int)bands.size(); }

    // Resolve one <SimpleSource>/<VRTRawRasterBand> into a RawBand.
    bool resolve_source_file(const std::string& file, RawBand& rb, int depth = 0) {
        if (depth > 4) { error = "VRT nesting too deep: " + file; return false; }
        std::string txt;
        const std::string hdr = envi_hdr_path(file);
        const bool looks_vrt = file.size() > 4 && lower(file.substr(file.size() - 4)) == ".vrt";
        if (looks_vrt) {
            if (!slurp(file, txt)) { error = "cannot read " + file; return false; }
            XmlElem ds;
            if (xml_next(txt, "VRTDataset", 0, ds) == std::string::npos) { error = "not a VRT: " + file; return false; }
            XmlElem band;
            if (xml_next(ds.body, "VRTRasterBand", 0, band) == std::string::npos) { error = "VRT without bands: " + file; return false; }
            const int w = std::atoi(xml_attr(ds.open_tag, "rasterXSize").c_str());
            const int hgt = std::atoi(xml_attr(ds.open_tag, "rasterYSize").c_str());
            return resolve_band(band, dirname_of(file), w, hgt, rb, depth + 1);
        }
        if (!hdr.empty()) {                              // ENVI single-band (BIP with 1 band == flat)
            EnviHeader h;
            if (!read_envi_header(hdr, h)) { error = "bad ENVI header " + hdr; return false; }
            if (h.bands != 1) { error = "multi-band ENVI source not supported as a stack band: " + file; return false; }
            rb.path = file; rb.image_offset = h.header_offset; rb.elem_bytes = envi_type_bytes(h.data_type);
            rb.pixel_offset = rb.elem_bytes; rb.line_offset = (long)rb.elem_bytes * h.samples;
            rb.src_width = h.samples; rb.src_height = h.lines;
            rb.dtype = (h.data_type == 6) ? "CFloat32" : (h.data_type == 1 ? "Byte" : "other");
            return true;
        }
        error = "unsupported source (need .vrt or ENVI): " + file;
        return false;
    }
    bool resolve_band(const XmlElem& band, const std::string& dir, int w, int hgt, RawBand& rb, int depth) {
        const std::string sub = xml_attr(band.open_tag, "subClass");
        std::string dt = xml_attr(band.open_tag, "dataType");
        if (sub == "VRTRawRasterBand") {
            std::string fn, tag, t;
            if (!xml_child_text(band.body, "SourceFilename", fn, &tag) && !xml_child_text(band.body, "sourceFilename", fn, &tag)) {
                error = "raw band without SourceFilename"; return false;
            }
            const bool rel = xml_attr(tag, "relativeToVRT") == "1" || xml_attr(tag, "relativetoVRT") == "1";
            rb.path = rel ? join_path(dir, fn) : fn;
            rb.dtype = dt.empty() ? "CFloat32" : dt;
            rb.elem_bytes = dtype_bytes(rb.dtype);
            rb.image_offset = xml_child_text(band.body, "ImageOffset", t) ? std::atol(t.c_str()) : 0;
            rb.pixel_offset = xml_child_text(band.body, "PixelOffset", t) ? std::atol(t.c_str()) : rb.elem_bytes;
            rb.line_offset = xml_child_text(band.body, "LineOffset", t) ? std::atol(t.c_str()) : (long)rb.pixel_offset * w;
            if (xml_child_text(band.body, "ByteOrder", t) && lower(t) != "lsb") { error = "only LSB raw bands supported"; return false; }
            rb.src_width = w; rb.src_height = hgt;
            return true;
        }
        XmlElem src;
        if (xml_next(band.body, "SimpleSource", 0, src) == std::string::npos) { error = "band without SimpleSource"; return false; }
        std::string fn, tag;
        if (!xml_child_text(src.body, "SourceFilename", fn, &tag)) { error = "SimpleSource without SourceFilename"; return false; }
        const bool rel = xml_attr(tag, "relativeToVRT") == "1";
        if (!resolve_source_file(rel ? join_path(dir, fn) : fn, rb, depth)) return false;
        XmlElem rect;
        if (xml_next(src.body, "SrcRect", 0, rect) != std::string::npos) {
            rb.x_off += std::atoi(xml_attr(rect.open_tag, "xOff").c_str());
            rb.y_off += std::atoi(xml_attr(rect.open_tag, "yOff").c_str());
        }
        return true;
    }

    bool open(const std::string& p) {
        if (open_builtin(p)) return true;
#ifdef FRINGE_WITH_GDAL
        const std::string why = error;
        close_all(); bands.clear(); interleaved = false;
        if (open_gdal(p)) return true;
        error = why + "; GDAL: " + error;
#endif
        return false;
    }

#ifdef FRINGE_WITH_GDAL
    static int envi_type_of(GDALDataType t) {
        switch (t) {
            case GDT_Byte: return 1; case GDT_Int16: return 2; case GDT_Int32: return 3; case GDT_Float32: return 4;
            case GDT_Float64: return 5; case GDT_CFloat32: return 6; case GDT_CFloat64: return 9; case GDT_UInt16: return 12;
            case GDT_UInt32: return 13; default: return 0;
        }
    }
    // Any raster GDAL can open.  A dataset whose bands are CFloat32 is presented as a band-per-date stack (with the
    // bands' "slc" metadata domain, evd.cpp:212-221 / nmap.cpp:204-233), anything else as a pixel-interleaved raster
    // (weights, masks), with the ENVI-domain items the drivers look at (HALFWINDOWX / HALFWINDOWY, nmap.cpp:588-591).
    bool open_gdal(const std::string& p) {
        static std::once_flag once;
        std::call_once(once, [] { GDALAllRegister(); });
        gdal_ds = GDALOpen(p.c_str(), GA_ReadOnly);
        if (!gdal_ds) { error = "GDALOpen failed for " + p; return false; }
        path = p;
        cols = GDALGetRasterXSize(gdal_ds); rows = GDALGetRasterYSize(gdal_ds);
        const int nb = GDALGetRasterCount(gdal_ds);
        if (cols <= 0 || rows <= 0 || nb <= 0) { error = "empty raster " + p; return false; }
        const GDALDataType t0 = GDALGetRasterDataType(GDALGetRasterBand(gdal_ds, 1));
        for (int b = 1; b <= nb; ++b) {
            GDALRasterBandH hb = GDALGetRasterBand(gdal_ds, b);
            RawBand rb;
            rb.path = p;
            rb.dtype = GDALGetDataTypeName(GDALGetRasterDataType(hb));
            rb.elem_bytes = GDALGetDataTypeSizeBytes(GDALGetRasterDataType(hb));
            rb.src_width = cols; rb.src_height = rows;
            if (char** md = GDALGetMetadata(hb, "slc"))
                for (char** it = md; *it; ++it) {
                    const std::string kv(*it);
                    const size_t eq = kv.find('=');
                    if (eq != std::string::npos) rb.md_slc[kv.substr(0, eq)] = kv.substr(eq + 1);
                }
            bands.push_back(rb);
        }
        interleaved = (t0 != GDT_CFloat32);
        if (interleaved) {
            envi.samples = cols; envi.lines = rows; envi.bands = nb; envi.data_type = envi_type_of(t0); envi.interleave = "bip";
            for (const char* key : {"HALFWINDOWX", "HALFWINDOWY"})
                if (const char* v = GDALGetMetadataItem(gdal_ds, key, "ENVI")) envi.fields[lower(key)] = v;
        }
        return true;
    }
    bool gdal_read(int band /*1-based, 0 = all bands pixel-interleaved*/, int yoff, int n, void* dst, GDALDataType as) {
        std::lock_guard<std::mutex> g(gdal_mu);
        const int sz = GDALGetDataTypeSizeBytes(as);
        CPLErr e;
        if (band > 0)
            e = GDALRasterIO(GDALGetRasterBand(gdal_ds, band), GF_Read, 0, yoff, cols, n, dst, cols, n, as, 0, 0);
        else {
            const int nb = GDALGetRasterCount(gdal_ds);
            e = GDALDatasetRasterIO(gdal_ds, GF_Read, 0, yoff, cols, n, dst, cols, n, as, nb, nullptr,
                                    (GSpacing)sz * nb, (GSpacing)sz * nb * cols, (GSpacing)sz);
        }
        if (e != CE_None) { error = "GDAL read failed on " + path; return false; }
        return true;
    }
#endif

    bool open_builtin(const std::string& p) {
        path = p;
        const bool looks_vrt = p.size() > 4 && lower(p.substr(p.size() - 4)) == ".vrt";
        if (looks_vrt) {
            std::string txt;
            if (!slurp(p, txt)) { error = "cannot read " + p; return false; }
            XmlElem ds;
            if (xml_next(txt, "VRTDataset", 0, ds) == std::string::npos) { error = "not a VRT: " + p; return false; }
            cols = std::atoi(xml_attr(ds.open_tag, "rasterXSize").c_str());
            rows = std::atoi(xml_attr(ds.open_tag, "rasterYSize").c_str());
            size_t pos = 0;
            XmlElem band;
            while ((pos = xml_next(ds.body, "VRTRasterBand", pos, band)) != std::string::npos) {
                RawBand rb;
                if (!resolve_band(band, dirname_of(p), cols, rows, rb, 0)) return false;
                // <Metadata domain="slc"><MDI key="Date">...</MDI></Metadata>
                size_t mp = 0;
                XmlElem md;
                while ((mp = xml_next(band.body, "Metadata", mp, md)) != std::string::npos) {
                    if (xml_attr(md.open_tag, "domain") != "slc") continue;
                    size_t ip = 0;
                    XmlElem it;
                    while ((ip = xml_next(md.body, "MDI", ip, it)) != std::string::npos)
                        rb.md_slc[xml_attr(it.open_tag, "key")] = trim(it.body);
                }
                rb.fd = ::open(rb.path.c_str(), O_RDONLY);
                if (rb.fd < 0) { error = "cannot open " + rb.path; return false; }
                bands.push_back(rb);
            }
            if (cols <= 0 || rows <= 0 || bands.empty()) { error = "empty VRT " + p; return false; }
            return true;
        }
        const std::string hdr = envi_hdr_path(p);
        if (hdr.empty() || !read_envi_header(hdr, envi)) { error = "cannot open raster (need .vrt or ENVI .hdr): " + p; return false; }
        interleaved = true;
        cols = envi.samples; rows = envi.lines;
        if (envi.bands > 1 && envi.interleave != "bip") { error = "only BIP multi-band ENVI supported: " + p; return false; }
        fd = ::open(p.c_str(), O_RDONLY);
        if (fd < 0) { error = "cannot open " + p; return false; }
        return true;
    }

    // Read lines [yoff, yoff+n) of band b (0-based) as tightly packed samples.
    bool read_band_lines(int b, int yoff, int n, void* dst, int elem_bytes) {
        if (b < 0 || b >= (int)bands.size()) { error = "band index outside the raster (not a band-per-file stack?)"; return false; }
        const RawBand& rb = bands[b];
#ifdef FRINGE_WITH_GDAL
        if (gdal_ds) {
            if (rb.elem_bytes != elem_bytes) { error = "unexpected sample type in " + rb.path; return false; }
            return gdal_read(b + 1, yoff, n, dst, GDALGetRasterDataType(GDALGetRasterBand(gdal_ds, b + 1)));
        }
#endif
        if (rb.elem_bytes != elem_bytes) { error = "unexpected sample type in " + rb.path; return false; }
        char* out = static_cast<char*>(dst);
        const bool packed = rb.pixel_offset == elem_bytes;
        // whole-width lines stored back to back: the block of this band is one contiguous byte range
        if (packed && rb.x_off == 0 && rb.line_offset == (long)cols * elem_bytes) {
            const long off = rb.image_offset + (long)(rb.y_off + yoff) * rb.line_offset;
            if (!pread_all(rb.fd, out, (size_t)n * cols * elem_bytes, off)) { error = "short read from " + rb.path; return false; }
            return true;
        }
        std::vector<char> tmp;
        for (int r = 0; r < n; ++r) {
            const long off = rb.image_offset + (long)(rb.y_off + yoff + r) * rb.line_offset + (long)rb.x_off * rb.pixel_offset;
            if (packed) {
                const size_t want = (size_t)cols * elem_bytes;
                if (!pread_all(rb.fd, out + (size_t)r * want, want, off)) { error = "short read from " + rb.path; return false; }
            } else {
                tmp.resize((size_t)cols * rb.pixel_offset);
                if (!pread_all(rb.fd, tmp.data(), tmp.size(), off)) { error = "short read from " + rb.path; return false; }
                for (int c = 0; c < cols; ++c) std::memcpy(out + ((size_t)r * cols + c) * elem_bytes, tmp.data() + (size_t)c * rb.pixel_offset, elem_bytes);
            }
        }
        return true;
    }
    // Read lines of an interleaved (BIP) ENVI file: all bands, pixel-interleaved, packed.
    bool read_interleaved_lines(int yoff, int n, void* dst) {
#ifdef FRINGE_WITH_GDAL
        if (gdal_ds) return gdal_read(0, yoff, n, dst, GDALGetRasterDataType(GDALGetRasterBand(gdal_ds, 1)));
#endif
        const size_t line_bytes = (size_t)cols * envi.bands * envi_type_bytes(envi.data_type);
        if (!pread_all(fd, dst, line_bytes * (size_t)n, (off_t)envi.header_offset + (off_t)yoff * (off_t)line_bytes)) { error = "short read from " + path; return false; }
        return true;
    }
    // Lines of a single-band raster of any integer / float type as a byte mask (non-zero -> 1), the
    // conversion GDAL's RasterIO(..., GDT_Byte) applies for nmap.cpp:323-343 in effect.
    bool read_mask_lines(int yoff, int n, uint8_t* dst, std::vector<char>& scratch) {
#ifdef FRINGE_WITH_GDAL
        if (gdal_ds) {                                  // GDAL converts to Byte exactly as nmap.cpp:323-343 asks it to
            if (!gdal_read(1, yoff, n, dst, GDT_Byte)) return false;
            return true;
        }
#endif
        int eb = 0; bool is_float = false;
        if (interleaved) { eb = envi_type_bytes(envi.data_type); is_float = (envi.data_type == 4 || envi.data_type == 5); }
        else if (!bands.empty()) { eb = bands[0].elem_bytes; is_float = lower(bands[0].dtype).find("float") != std::string::npos; }
        const size_t npx = (size_t)cols * n;
        if (eb == 1) return interleaved ? read_interleaved_lines(yoff, n, dst) : read_band_lines(0, yoff, n, dst, 1);
        if (eb != 2 && eb != 4 && eb != 8) { error = "unsupported mask sample type in " + path; return false; }
        scratch.resize(npx * eb);
        if (!(interleaved ? read_interleaved_lines(yoff, n, scratch.data()) : read_band_lines(0, yoff, n, scratch.data(), eb))) return false;
        for (size_t k = 0; k < npx; ++k) {
            bool nz;
            if (eb == 2) { uint16_t v; std::memcpy(&v, scratch.data() + 2 * k, 2); nz = v != 0; }
            else if (eb == 4 && is_float) { float v; std::memcpy(&v, scratch.data() + 4 * k, 4); nz = v != 0.f; }
            else if (eb == 4) { uint32_t v; std::memcpy(&v, scratch.data() + 4 * k, 4); nz = v != 0; }
            else { double v; std::memcpy(&v, scratch.data() + 8 * k, 8); nz = v != 0.0; }
            dst[k] = nz ? 1 : 0;
        }
        return true;
    }
    // true when every band is a flat CFloat32 file (what the drivers read as the SLC stack)
    bool is_cfloat32_stack() const {
        if (interleaved || bands.empty()) return false;
        for (const auto& b : bands) if (lower(b.dtype) != "cfloat32" || b.elem_bytes != 8) return false;
        return true;
    }
};

// ---- ENVI writer (INTERLEAVE=BIP, SUFFIX=ADD) ------------------------------------------------------
struct EnviWriter {
    std::string path;
    int cols = 0, rows = 0, bands = 1, data_type = 4, fd = -1;
    std::vector<std::pair<std::string, std::string>> extra;
    bool create(const std::string& p, int c, int r, int nb, int envi_type) {
        path = p; cols = c; rows = r; bands = nb; data_type = envi_type;
        fd = ::open(p.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0666);
        if (fd < 0) return false;
        const off_t total = (off_t)cols * (off_t)rows * (off_t)bands * (off_t)envi_type_bytes(envi_type);
        if (::ftruncate(fd, total) != 0) return false;
        return write_header();
    }
    void set_metadata(const std::string& k, const std::string& v) { extra.emplace_back(k, v); }
    bool write_header() const {
        std::ofstream h((path + ".hdr").c_str());
        if (!h) return false;
        h << "ENVI\ndescription = {\n" << path << "}\nsamples = " << cols << "\nlines   = " << rows << "\nbands   = " << bands
          << "\nheader offset = 0\nfile type = ENVI Standard\ndata type = " << data_type
          << "\ninterleave = bip\nbyte order = 0\n";
        for (auto& kv : extra) h << kv.first << " = " << kv.second << "\n";
        return (bool)h;
    }
    bool write_lines(int y0, int n, const void* src) {
        const size_t line_bytes = (size_t)cols * bands * envi_type_bytes(data_type);
        return pwrite_all(fd, src, line_bytes * (size_t)n, (off_t)y0 * (off_t)line_bytes);
    }
    bool close_file() {
        bool ok = true;
        if (fd >= 0) { ok = write_header(); ::close(fd); fd = -1; }
        return ok;
    }
    ~EnviWriter() { if (fd >= 0) ::close(fd); }
};

}  // namespace fringe_host
