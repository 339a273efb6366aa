// Reader for Kiwi's HDF5 Green's function databases: <base>.index + <base>.<i>.chunk as written by
// gfdb_io_hdf.f90 (:119-180 index file of scalar datasets; :236-310 chunk file with the "index" dataset of object
// references; :313-427 one 1-D float dataset per trace under /gf/<ixc>/<iz>/<ig> with the integer attributes "pofs"
// and "ofs"; :429-524 the read side this file replaces; trace_from_storable sparse_trace.f90:849-878).
//
// libhdf5 is not available to this library, so this is a minimal parser of the HDF5 file format as published by
// The HDF Group ("HDF5 File Format Specification", versions 1.0/1.1 of the structures HDF5 1.6/1.8 write by default:
// superblock 0/1, version-1 object headers, symbol-table groups with version-1 B-trees and local heaps, contiguous or
// compact dataset layout, version 1-3 attribute messages, object references = addresses of object headers).
// Everything else (superblock 2/3, "OHDR" object headers, chunked/filtered layout, external storage) is rejected with
// a message.  Pinning: the generic structures (superblock behind a user block, base address, root symbol table, B-tree, heap,
// version-1 object header, dataspace / datatype / layout / attribute messages) are checked on the one file in this image that the
// real library wrote (scipy's MATLAB 7.3 test file, HDF5 1.6); Kiwi's own layout on top of them (index of object references, one
// dataset with two attributes per trace) only on files of an independent minimal writer that follows the same specification
// (tests/h5mini_writer.py) -- no Kiwi database and no libhdf5 exist here.
#include "kiwi_internal.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace {

const uint64_t UNDEF = ~0ull;

struct H5Error { std::string msg; };
[[noreturn]] void fail(const std::string& m) { throw H5Error{m}; }

struct Dataset {
    int dtype_class = -1;            // 0 fixed point, 1 floating point, 7 reference
    uint32_t dtype_size = 0;
    std::vector<uint64_t> dims;      // empty = scalar
    const uint8_t* data = nullptr;   // raw little-endian elements (contiguous or compact), null = never written
    uint64_t nbytes = 0;
    struct Attr { std::string name; int dtype_class; uint32_t dtype_size; std::vector<uint64_t> dims; const uint8_t* data; };
    std::vector<Attr> attrs;
    uint64_t nelem() const {   // (saturating: a product that overflows is larger than any file)
        uint64_t n = 1;
        for (uint64_t d : dims) { if (d != 0 && n > (1ull << 56) / d) return 1ull << 56; n *= d; }
        return n;
    }
};

class H5File {
  public:
    explicit H5File(const std::string& path) : path_(path) {
        fd_ = open(path.c_str(), O_RDONLY);
        if (fd_ < 0) fail("gfdb: failed to open file: " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 64) { close(fd_); fail("gfdb: not an HDF5 file: " + path); }
        size_ = (uint64_t)st.st_size;
        map_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (map_ == MAP_FAILED) { close(fd_); fail("gfdb: cannot map file: " + path); }
        try { read_superblock(); } catch (...) { munmap((void*)map_, size_); close(fd_); throw; }
    }
    ~H5File() { if (map_ && map_ != MAP_FAILED) munmap((void*)map_, size_); if (fd_ >= 0) close(fd_); }
    H5File(const H5File&) = delete;
    H5File& operator=(const H5File&) = delete;

    const std::map<std::string, uint64_t>& root_members() const { return root_; }
    // an object reference as stored in a dataset (address relative to the file's base address) -> object header address
    uint64_t deref(uint64_t stored) const { return stored + base_; }
    // object header address of a member of the root group, UNDEF if absent
    uint64_t root_member(const std::string& name) const {
        auto it = root_.find(name);
        return it == root_.end() ? UNDEF : it->second;
    }
    Dataset dataset(uint64_t ohdr_addr) const {
        Dataset d;
        bool have_layout = false;
        for_each_message(ohdr_addr, [&](unsigned type, const uint8_t* p, uint64_t n) {
            if (type == 0x0001) d.dims = parse_dataspace(p, n);
            else if (type == 0x0003) parse_datatype(p, n, &d.dtype_class, &d.dtype_size);
            else if (type == 0x0008) { parse_layout(p, n, &d.data, &d.nbytes); have_layout = true; }
            else if (type == 0x000C) d.attrs.push_back(parse_attribute(p, n));
        });
        if (d.dtype_class < 0 || !have_layout) fail("gfdb: object is not a dataset in file: " + path_);
        if (d.data && d.nbytes < d.nelem() * d.dtype_size) fail("gfdb: dataset storage shorter than its extent in file: " + path_);
        return d;
    }

  private:
    std::string path_;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    uint64_t size_ = 0, base_ = 0, sb_ = 0;
    unsigned O_ = 8, L_ = 8;
    std::map<std::string, uint64_t> root_;

    const uint8_t* at(uint64_t off, uint64_t n) const {
        if (off > size_ || n > size_ - off) fail("gfdb: address outside of file (truncated or not a supported HDF5 layout): " + path_);
        return map_ + off;
    }
    static uint64_t le(const uint8_t* p, unsigned n) { uint64_t v = 0; for (unsigned i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
    uint64_t addr_at(const uint8_t* p) const {   // file address field: all ones = undefined
        const uint64_t v = le(p, O_);
        const uint64_t undef = O_ == 8 ? ~0ull : ((1ull << (8 * O_)) - 1);
        return v == undef ? UNDEF : v + base_;
    }

    void read_superblock() {
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t off = 0;
        bool found = false;
        for (; off + 8 <= size_; off = off ? off * 2 : 512) {   // 0, 512, 1024, ...
            if (memcmp(map_ + off, sig, 8) == 0) { found = true; break; }
            if (off > (1u << 24)) break;
        }
        if (!found) fail("gfdb: not an HDF5 file: " + path_);
        sb_ = off;
        const uint8_t* p = at(off, 24);
        const unsigned ver = p[8];
        if (ver > 1) fail("gfdb: HDF5 superblock version " + std::to_string(ver) + " is not supported (files of HDF5 1.6/1.8 defaults are): " + path_);
        O_ = p[13]; L_ = p[14];
        if ((O_ != 4 && O_ != 8) || (L_ != 4 && L_ != 8)) fail("gfdb: unsupported size of offsets/lengths in file: " + path_);
        uint64_t q = off + 24 + (ver == 1 ? 4 : 0);   // v1: indexed storage internal node K + reserved
        base_ = le(at(q, O_), O_); q += 4 * O_;     // base, free-space info, end of file, driver info
        base_ += 0;
        // root group symbol table entry
        const uint8_t* e = at(q, 2 * O_ + 24);
        const uint64_t root_ohdr = addr_at(e + O_);
        const unsigned cache = (unsigned)le(e + 2 * O_, 4);
        uint64_t btree = UNDEF, heap = UNDEF;
        if (cache == 1) { btree = addr_at(e + 2 * O_ + 8); heap = addr_at(e + 2 * O_ + 8 + O_); }
        else group_of(root_ohdr, &btree, &heap);
        if (btree == UNDEF || heap == UNDEF) fail("gfdb: root group without symbol table in file: " + path_);
        list_group(btree, heap, &root_);
    }

    void group_of(uint64_t ohdr, uint64_t* btree, uint64_t* heap) const {
        for_each_message(ohdr, [&](unsigned type, const uint8_t* p, uint64_t n) {
            if (type == 0x0011 && n >= 2 * O_) { *btree = addr_at(p); *heap = addr_at(p + O_); }
        });
    }

    void list_group(uint64_t btree, uint64_t heap, std::map<std::string, uint64_t>* out) const {
        const uint8_t* h = at(heap, 8 + 2 * L_ + O_);
        if (memcmp(h, "HEAP", 4) != 0) fail("gfdb: bad local heap signature in file: " + path_);
        const uint64_t hsize = le(h + 8, L_), hdata = addr_at(h + 8 + 2 * L_);
        walk_btree(btree, hdata, hsize, out, 0);
    }
    void walk_btree(uint64_t node, uint64_t hdata, uint64_t hsize, std::map<std::string, uint64_t>* out, int depth) const {
        if (depth > 32) fail("gfdb: group B-tree too deep in file: " + path_);
        const uint8_t* p = at(node, 8 + 2 * O_);
        if (memcmp(p, "TREE", 4) != 0) fail("gfdb: bad B-tree signature in file: " + path_);
        if (p[4] != 0) fail("gfdb: unexpected B-tree type in a group in file: " + path_);
        const unsigned level = p[5], n = (unsigned)le(p + 6, 2);
        const uint8_t* kc = at(node + 8 + 2 * O_, (uint64_t)n * (L_ + O_) + L_);
        for (unsigned i = 0; i < n; i++) {
            const uint64_t child = addr_at(kc + L_ + (uint64_t)i * (L_ + O_));
            if (level > 0) { walk_btree(child, hdata, hsize, out, depth + 1); continue; }
            const uint8_t* s = at(child, 8);
            if (memcmp(s, "SNOD", 4) != 0) fail("gfdb: bad symbol table node signature in file: " + path_);
            const unsigned nsym = (unsigned)le(s + 6, 2);
            const uint64_t esz = 2 * O_ + 24;
            const uint8_t* e = at(child + 8, nsym * esz);
            for (unsigned k = 0; k < nsym; k++, e += esz) {
                const uint64_t noff = le(e, O_);
                if (noff >= hsize) fail("gfdb: link name outside of the local heap in file: " + path_);
                const char* nm = (const char*)at(hdata + noff, 1);
                const size_t maxlen = (size_t)(hsize - noff);
                (*out)[std::string(nm, strnlen(nm, maxlen))] = addr_at(e + O_);
            }
        }
    }

    template <class F>
    void for_each_message(uint64_t ohdr, F f) const {
        const uint8_t* p = at(ohdr, 16);
        if (memcmp(p, "OHDR", 4) == 0) fail("gfdb: version 2 object headers are not supported (file written with a 'latest' format setting): " + path_);
        if (p[0] != 1) fail("gfdb: unexpected object header version in file: " + path_);
        unsigned nmsgs = (unsigned)le(p + 2, 2);
        std::vector<std::pair<uint64_t, uint64_t>> blocks;   // (address, length) of message blocks
        blocks.push_back({ohdr + 16, le(p + 8, 4)});
        for (size_t b = 0; b < blocks.size() && nmsgs > 0; b++) {
            uint64_t q = blocks[b].first;
            const uint64_t end = q + blocks[b].second;
            while (q + 8 <= end && nmsgs > 0) {
                const uint8_t* m = at(q, 8);
                const unsigned type = (unsigned)le(m, 2);
                const uint64_t n = le(m + 2, 2);
                const uint8_t* body = at(q + 8, n);
                nmsgs--;
                if (type == 0x0010) {   // continuation
                    if (n < O_ + L_) fail("gfdb: bad continuation message in file: " + path_);
                    blocks.push_back({addr_at(body), le(body + O_, L_)});
                } else f(type, body, n);
                q += 8 + n;
            }
        }
    }

    std::vector<uint64_t> parse_dataspace(const uint8_t* p, uint64_t n) const {
        if (n < 4) fail("gfdb: bad dataspace message in file: " + path_);
        const unsigned ver = p[0], rank = p[1];
        uint64_t q;
        if (ver == 1) q = 8;
        else if (ver == 2) q = 4;
        else fail("gfdb: unsupported dataspace message version in file: " + path_);
        if (n < q + (uint64_t)rank * L_) fail("gfdb: bad dataspace message in file: " + path_);
        std::vector<uint64_t> dims(rank);
        for (unsigned i = 0; i < rank; i++) dims[i] = le(p + q + (uint64_t)i * L_, L_);
        return dims;
    }
    void parse_datatype(const uint8_t* p, uint64_t n, int* cls, uint32_t* size) const {
        if (n < 8) fail("gfdb: bad datatype message in file: " + path_);
        *cls = p[0] & 0x0f;
        *size = (uint32_t)le(p + 4, 4);
        if ((*cls == 0 || *cls == 1) && (p[1] & 1)) fail("gfdb: big-endian data is not supported: " + path_);
    }
    // size in bytes of a datatype message (needed to step over it inside a version-2/3 attribute)
    void parse_layout(const uint8_t* p, uint64_t n, const uint8_t** data, uint64_t* nbytes) const {
        const std::string bad = "gfdb: bad layout message in file: " + path_;
        if (n < 3) fail(bad);
        const unsigned ver = p[0];
        if (ver == 3) {
            const unsigned cls = p[1];
            if (cls == 0) {
                if (n < 4) fail(bad);
                const uint64_t sz = le(p + 2, 2);
                if (n < 4 + sz) fail("gfdb: bad compact layout in file: " + path_);
                *data = p + 4; *nbytes = sz;
            } else if (cls == 1) {
                if (n < 2 + (uint64_t)O_ + L_) fail(bad);
                const uint64_t a = addr_at(p + 2), sz = le(p + 2 + O_, L_);
                *nbytes = sz; *data = a == UNDEF ? nullptr : at(a, sz);
            } else fail("gfdb: chunked dataset layout is not supported (Kiwi writes contiguous datasets): " + path_);
        } else if (ver == 1 || ver == 2) {
            const unsigned rank = p[1], cls = p[2];
            if (rank > 32) fail(bad);
            if (cls == 1) {
                if (n < 8 + (uint64_t)O_ + 4 * (uint64_t)rank) fail(bad);
                const uint64_t a = addr_at(p + 8);
                uint64_t sz = 1;   // dimension sizes follow the address (4 bytes each); the last one is the element size
                for (unsigned i = 0; i < rank; i++) sz = mul_checked(sz, le(p + 8 + O_ + 4 * i, 4));
                *nbytes = sz; *data = a == UNDEF ? nullptr : at(a, sz);
            } else if (cls == 0) {
                const uint64_t q = 8 + 4 * (uint64_t)rank;
                if (n < q + 4) fail(bad);
                const uint64_t sz = le(p + q, 4);
                if (n < q + 4 + sz) fail("gfdb: bad compact layout in file: " + path_);
                *data = p + q + 4; *nbytes = sz;
            } else fail("gfdb: chunked dataset layout is not supported (Kiwi writes contiguous datasets): " + path_);
        } else fail("gfdb: unsupported layout message version in file: " + path_);
    }
    // products of file-controlled sizes: refuse what cannot be a size inside the file
    uint64_t mul_checked(uint64_t a, uint64_t b) const {
        if (a != 0 && b > (1ull << 48) / a) fail("gfdb: implausible dataset size in file: " + path_);
        return a * b;
    }
    static uint64_t pad8(uint64_t v) { return (v + 7) & ~7ull; }
    Dataset::Attr parse_attribute(const uint8_t* p, uint64_t n) const {
        if (n < 8) fail("gfdb: bad attribute message in file: " + path_);
        const unsigned ver = p[0];
        const uint64_t nsz = le(p + 2, 2), tsz = le(p + 4, 2), ssz = le(p + 6, 2);
        uint64_t q = ver == 3 ? 9 : 8;
        if (ver < 1 || ver > 3) fail("gfdb: unsupported attribute message version in file: " + path_);
        const bool padded = ver == 1;
        Dataset::Attr a;
        if (q + nsz > n) fail("gfdb: bad attribute message in file: " + path_);
        a.name = std::string((const char*)p + q, strnlen((const char*)p + q, (size_t)nsz));
        q += padded ? pad8(nsz) : nsz;
        if (q + tsz > n) fail("gfdb: bad attribute message in file: " + path_);
        parse_datatype(p + q, tsz, &a.dtype_class, &a.dtype_size);
        q += padded ? pad8(tsz) : tsz;
        if (q + ssz > n) fail("gfdb: bad attribute message in file: " + path_);
        a.dims = parse_dataspace(p + q, ssz);
        q += padded ? pad8(ssz) : ssz;
        uint64_t ne = 1;
        for (uint64_t d : a.dims) ne = mul_checked(ne, d);
        if (q + mul_checked(ne, a.dtype_size) > n) fail("gfdb: attribute data outside of its message in file: " + path_);
        a.data = p + q;
        return a;
    }
};

float scalar_real(const H5File& f, const char* name, const std::string& path, bool optional = false, float dflt = 0.f) {
    const uint64_t a = f.root_member(name);
    if (a == UNDEF) { if (optional) return dflt; fail("gfdb: failed to read dataset from file: " + path); }
    const Dataset d = f.dataset(a);
    if (d.dtype_class != 1 || d.dtype_size != 4 || d.nelem() != 1 || !d.data) fail("gfdb: failed to read dataset from file: " + path);
    float v; memcpy(&v, d.data, 4); return v;
}
int scalar_int(const H5File& f, const char* name, const std::string& path) {
    const uint64_t a = f.root_member(name);
    if (a == UNDEF) fail("gfdb: failed to read dataset from file: " + path);
    const Dataset d = f.dataset(a);
    if (d.dtype_class != 0 || d.dtype_size != 4 || d.nelem() != 1 || !d.data) fail("gfdb: failed to read dataset from file: " + path);
    int32_t v; memcpy(&v, d.data, 4); return v;
}

}  // namespace

extern "C" kiwi_gfdb* kiwi_gfdb_read_hdf(const char* basepath) {
    if (!basepath) { kiwi_set_error("kiwi_gfdb_read_hdf: null path"); return nullptr; }
    kiwi_gfdb* db = nullptr;
    try {
        const std::string base(basepath), ipath = base + ".index";
        float dt, dx, dz, firstx, firstz;
        int nchunks, nx, nxc, nz, ng;
        {   // gfdb_io_read_index, gfdb_io_hdf.f90:119-180 (firstx, firstz are optional: older databases lack them)
            H5File f(ipath);
            dt = scalar_real(f, "dt", ipath); dx = scalar_real(f, "dx", ipath); dz = scalar_real(f, "dz", ipath);
            firstx = scalar_real(f, "firstx", ipath, true, 0.f); firstz = scalar_real(f, "firstz", ipath, true, 0.f);
            nchunks = scalar_int(f, "nchunks", ipath); nx = scalar_int(f, "nx", ipath); nxc = scalar_int(f, "nxc", ipath);
            nz = scalar_int(f, "nz", ipath); ng = scalar_int(f, "ng", ipath);
        }
        if (nchunks < 1 || nxc < 1 || nx < 1 || (long long)nxc * (nchunks - 1) >= nx) fail("gfdb: inconsistent chunk layout in file: " + ipath);
        db = kiwi_gfdb_create(nx, nz, ng, dt, dx, dz, firstx, firstz);
        if (!db) return nullptr;
        std::vector<float> dense;
        for (int ichunk = 1; ichunk <= nchunks; ichunk++) {
            const std::string cpath = base + "." + std::to_string(ichunk) + ".chunk";
            const int nxcthis = ichunk == nchunks ? nx - (ichunk - 1) * nxc : nxc;   // gfdb.f90:251-253
            H5File f(cpath);
            const uint64_t ia = f.root_member("index");
            if (ia == UNDEF) fail("gfdb: failed to open index dataset: " + cpath);
            const Dataset idx = f.dataset(ia);
            // Fortran dims (ng, nz, nxc) = C order (nxc, nz, ng): ig runs fastest
            if (idx.dtype_class != 7 || idx.dtype_size != 8 || idx.dims.size() != 3 || (int)idx.dims[0] != nxcthis || (int)idx.dims[1] != nz ||
                (int)idx.dims[2] != ng || !idx.data)
                fail("gfdb: failed to read index dataset: " + cpath);
            for (int ixc = 1; ixc <= nxcthis; ixc++)
                for (int iz = 1; iz <= nz; iz++)
                    for (int ig = 1; ig <= ng; ig++) {
                        uint64_t ref;
                        memcpy(&ref, idx.data + 8 * (((size_t)(ixc - 1) * nz + (iz - 1)) * ng + (ig - 1)), 8);
                        if (ref == 0) continue;   // no trace stored (gfdb.f90: references initialised to 0)
                        const Dataset tr = f.dataset(f.deref(ref));   // h5rdereference: the reference is the object header's address
                        if (tr.dtype_class != 1 || tr.dtype_size != 4 || tr.dims.size() != 1 || !tr.data) fail("gfdb: failed to get a dataset: " + cpath);
                        const Dataset::Attr *pofs = nullptr, *ofs = nullptr;
                        // the reference reads attribute 0 as pofs and attribute 1 as ofs (creation order, :451-470); names are checked here
                        for (const Dataset::Attr& a : tr.attrs) { if (a.name == "pofs") pofs = &a; else if (a.name == "ofs") ofs = &a; }
                        if (!pofs || !ofs || pofs->dtype_class != 0 || ofs->dtype_class != 0 || pofs->dtype_size != 4 || ofs->dtype_size != 4 ||
                            pofs->dims.size() != 1 || ofs->dims != pofs->dims || pofs->dims[0] < 1 || pofs->dims[0] > (1u << 24) ||
                            tr.dims[0] > (1ull << 28))
                            fail("gfdb: failed to get attributes of a dataset: " + cpath);
                        const int nstrips = (int)pofs->dims[0];
                        const long long npacked = (long long)tr.dims[0];
                        // trace_from_storable (sparse_trace.f90:849-878): strip i = packed[pofs(i) .. pofs(i+1)-1] starting at sample ofs(i);
                        // between the strips the trace is zero (sparse_trace.f90:29-50)
                        std::vector<int32_t> po(nstrips), of(nstrips);
                        memcpy(po.data(), pofs->data, 4 * (size_t)nstrips); memcpy(of.data(), ofs->data, 4 * (size_t)nstrips);
                        long long last = 0;
                        for (int s = 0; s < nstrips; s++) {
                            const long long n = (s + 1 < nstrips ? po[s + 1] : npacked + 1) - po[s];
                            if (po[s] < 1 || n < 1 || po[s] - 1 + n > npacked || (s > 0 && of[s] < last)) fail("gfdb: inconsistent strip offsets of a trace in: " + cpath);
                            last = (long long)of[s] + n;
                        }
                        const long long len = last - of[0];
                        if (len < 1 || len > (1LL << 28)) fail("gfdb: unreasonable trace length in: " + cpath);
                        dense.assign((size_t)len, 0.f);
                        for (int s = 0; s < nstrips; s++) {
                            const long long n = (s + 1 < nstrips ? po[s + 1] : npacked + 1) - po[s];
                            memcpy(dense.data() + (of[s] - of[0]), tr.data + 4 * (size_t)(po[s] - 1), 4 * (size_t)n);
                        }
                        const int ix = (ichunk - 1) * nxc + ixc;
                        if (kiwi_gfdb_save_array(db, ix, iz, ig, of[0], (int)len, dense.data())) { kiwi_gfdb_destroy(db); return nullptr; }
                    }
        }
        return db;
    } catch (const H5Error& e) {
        if (db) kiwi_gfdb_destroy(db);
        kiwi_set_error("%s", e.msg.c_str());
        return nullptr;
    } catch (const std::exception& e) {   // bad_alloc / length_error on sizes a corrupt file asks for: an error, not std::terminate
        if (db) kiwi_gfdb_destroy(db);
        kiwi_set_error("gfdb: %s while reading %s", e.what(), basepath);
        return nullptr;
    }
}

// Generic access to a dataset in the root group of an HDF5 file (what the index file of a database consists of; also how
// the parser is checked against a file written by the real library).  name == NULL: `buf` receives the NUL-separated member
// names of the root group.  Returns 0 on success; *nbytes = size of the data (or of the name list) even when cap is too small.
extern "C" int kiwi_h5_read_root_dataset(const char* path, const char* name, int* dtype_class, int* dtype_size, int* rank, long long* dims8,
                                          void* buf, long long cap, long long* nbytes, int* nattrs) {
    if (!path) return kiwi_set_error("kiwi_h5_read_root_dataset: null path");
    try {
        H5File f(path);
        if (!name) {
            std::string all;
            for (const auto& kv : f.root_members()) { all += kv.first; all.push_back('\0'); }
            if (nbytes) *nbytes = (long long)all.size();
            if (buf && cap > 0) memcpy(buf, all.data(), (size_t)std::min<long long>(cap, (long long)all.size()));
            return 0;
        }
        const uint64_t a = f.root_member(name);
        if (a == UNDEF) return kiwi_set_error("no object '%s' in the root group of %s", name, path);
        const Dataset d = f.dataset(a);
        if (dtype_class) *dtype_class = d.dtype_class;
        if (dtype_size) *dtype_size = (int)d.dtype_size;
        if (rank) *rank = (int)d.dims.size();
        if (dims8) for (size_t i = 0; i < d.dims.size() && i < 8; i++) dims8[i] = (long long)d.dims[i];
        if (nattrs) *nattrs = (int)d.attrs.size();
        const long long n = d.data ? (long long)(d.nelem() * d.dtype_size) : 0;
        if (nbytes) *nbytes = n;
        if (buf && cap > 0 && n > 0) memcpy(buf, d.data, (size_t)std::min(cap, n));
        return 0;
    } catch (const H5Error& e) {
        return kiwi_set_error("%s", e.msg.c_str());
    } catch (const std::exception& e) {
        return kiwi_set_error("%s while reading %s", e.what(), path);
    }
}
