// Memory-mapped reader of the reference's on-disk graph format: one uncompressed .npz per graph
// with the six SparseGraph members X, Ri_rows, Ri_cols, Ro_rows, Ro_cols, y, as written by
// save_graph (np.savez(filename, **graph._asdict()), gnn/graph.py:179-181) and read back by
// load_graph (gnn/graph.py:188-191).  Replaces np.load + the per-member array construction in front
// of the batch packer: the returned pointers point into the mapping (zero copy), so a batch of
// files goes from the page cache to the pinned staging buffers with one pass over the bytes.
//
// .npz = a ZIP archive whose members are .npy files, method "stored" (0) for np.savez.  numpy
// writes members with force_zip64, so sizes and offsets are taken from the CENTRAL directory
// (with its zip64 extra field when a 32-bit field is 0xFFFFFFFF), never from the local headers.
// .npy = "\x93NUMPY", version (1.0 / 2.0 / 3.0), header length, a Python dict literal with
// 'descr', 'fortran_order', 'shape', padded so that the data starts on a 64-byte boundary of the
// member (not of the archive: a member whose data is not aligned for its type is copied once).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gnnseg.h"

namespace {

struct Mapping {
    void* base = nullptr;
    size_t size = 0;
    std::vector<void*> copies;      // members that had to be re-aligned
};

inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

struct Member {
    const unsigned char* data = nullptr;   // first byte of the array data
    uint64_t bytes = 0;                    // bytes of array data
    char kind = 0;                         // 'f' or 'i'
    int item = 0;                          // item size in bytes
    int ndim = 0;
    int64_t shape[2] = {0, 0};
};

// Parse the header of a .npy member occupying [p, p + size).  Returns false on anything this
// reader does not handle (big-endian, Fortran order, more than 2 dimensions, object arrays).
bool parse_npy(const unsigned char* p, uint64_t size, Member* m) {
    if (size < 10 || std::memcmp(p, "\x93NUMPY", 6) != 0) return false;
    const int major = p[6];
    uint64_t hlen, hoff;
    if (major == 1) { hlen = rd16(p + 8); hoff = 10; }
    else if (major == 2 || major == 3) { if (size < 12) return false; hlen = rd32(p + 8); hoff = 12; }
    else return false;
    if (hoff + hlen > size) return false;
    const std::string h(reinterpret_cast<const char*>(p + hoff), hlen);
    // 'descr': '<f4'
    size_t k = h.find("'descr'");
    if (k == std::string::npos) return false;
    const size_t dcolon = h.find(':', k);
    if (dcolon == std::string::npos) return false;
    k = h.find('\'', dcolon);
    if (k == std::string::npos || k + 4 > h.size()) return false;
    const char order = h[k + 1];
    if (order != '<' && order != '|' && order != '=') return false;       // little endian only
    m->kind = h[k + 2];
    m->item = std::atoi(h.c_str() + k + 3);
    if ((m->kind != 'f' && m->kind != 'i' && m->kind != 'u') || (m->item != 1 && m->item != 4 && m->item != 8)) return false;
    // 'fortran_order': False
    k = h.find("'fortran_order'");
    if (k == std::string::npos) return false;
    const size_t colon = h.find(':', k);
    if (colon == std::string::npos) return false;
    const size_t val = h.find_first_not_of(' ', colon + 1);
    if (val == std::string::npos || val + 5 > h.size() || h.compare(val, 5, "False") != 0) return false;
    // 'shape': (N,) | (N, F) | ()
    k = h.find("'shape'");
    if (k == std::string::npos) return false;
    size_t a = h.find('(', k), b = h.find(')', k);
    if (a == std::string::npos || b == std::string::npos || b < a) return false;
    m->ndim = 0;
    size_t q = a + 1;
    while (q < b) {
        while (q < b && (h[q] == ' ' || h[q] == ',')) ++q;
        if (q >= b) break;
        if (m->ndim == 2) return false;
        char* end = nullptr;
        m->shape[m->ndim++] = std::strtoll(h.c_str() + q, &end, 10);
        q = (size_t)(end - h.c_str());
    }
    uint64_t count = 1;
    for (int i = 0; i < m->ndim; ++i) {
        if (m->shape[i] < 0) return false;
        if (m->shape[i] != 0 && count > size / (uint64_t)m->shape[i]) return false;     // more elements than the member has bytes
        count *= (uint64_t)m->shape[i];
    }
    if (count > size / (uint64_t)m->item) return false;
    m->data = p + hoff + hlen;
    m->bytes = count * (uint64_t)m->item;
    return hoff + hlen + m->bytes <= size;
}

// Walk the central directory of the archive mapped at [base, base + size): name -> stored member.
struct Entry { std::string name; uint64_t offset, csize, usize; int method; };

bool read_directory(const unsigned char* base, uint64_t size, std::vector<Entry>* out) {
    if (size < 22) return false;
    // end of central directory record: scan back over a possible archive comment
    int64_t eocd = -1;
    const uint64_t lo = size > 22 + 65535 ? size - 22 - 65535 : 0;
    for (uint64_t i = size - 22 + 1; i-- > lo;)
        if (rd32(base + i) == 0x06054b50u) { eocd = (int64_t)i; break; }
    if (eocd < 0) return false;
    uint64_t n = rd16(base + eocd + 10), cd_size = rd32(base + eocd + 12), cd_off = rd32(base + eocd + 16);
    if (n == 0xFFFF || cd_size == 0xFFFFFFFFu || cd_off == 0xFFFFFFFFu) {      // zip64: locator right before
        if (eocd < 20 || rd32(base + eocd - 20) != 0x07064b50u) return false;
        const uint64_t z = rd64(base + eocd - 20 + 8);
        if (z + 56 > size || rd32(base + z) != 0x06064b50u) return false;
        n = rd64(base + z + 32);
        cd_size = rd64(base + z + 40);
        cd_off = rd64(base + z + 48);
    }
    if (cd_off + cd_size > size) return false;
    uint64_t p = cd_off;
    for (uint64_t i = 0; i < n; ++i) {
        if (p + 46 > size || rd32(base + p) != 0x02014b50u) return false;
        Entry e;
        e.method = rd16(base + p + 10);
        e.csize = rd32(base + p + 20);
        e.usize = rd32(base + p + 24);
        const uint32_t nlen = rd16(base + p + 28), xlen = rd16(base + p + 30), clen = rd16(base + p + 32);
        e.offset = rd32(base + p + 42);
        if (p + 46 + nlen + xlen + clen > size) return false;
        e.name.assign(reinterpret_cast<const char*>(base + p + 46), nlen);
        // zip64 extended information (id 0x0001): the 64-bit values of the fields that are 0xFFFFFFFF, in order
        uint64_t x = p + 46 + nlen;
        const uint64_t xend = x + xlen;
        while (x + 4 <= xend) {
            const uint32_t id = rd16(base + x), len = rd16(base + x + 2);
            if (id == 0x0001) {
                uint64_t y = x + 4;
                if (e.usize == 0xFFFFFFFFu && y + 8 <= xend) { e.usize = rd64(base + y); y += 8; }
                if (e.csize == 0xFFFFFFFFu && y + 8 <= xend) { e.csize = rd64(base + y); y += 8; }
                if (e.offset == 0xFFFFFFFFu && y + 8 <= xend) { e.offset = rd64(base + y); y += 8; }
            }
            x += 4 + len;
        }
        out->push_back(e);
        p += 46 + nlen + xlen + clen;
    }
    return true;
}

}  // namespace

extern "C" int gnnseg_npz_open_graph_host(const char* path, GnnsegNpzGraph* g) {
    if (!path || !g) return GNNSEG_EINVAL;
    std::memset(g, 0, sizeof(*g));
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return GNNSEG_EIO;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { ::close(fd); return GNNSEG_EIO; }
    void* base = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) return GNNSEG_EIO;
    Mapping* map = new Mapping;
    map->base = base;
    map->size = (size_t)st.st_size;
    g->handle = map;
    const unsigned char* b = static_cast<const unsigned char*>(base);
    std::vector<Entry> dir;
    int rc = GNNSEG_OK;
    if (!read_directory(b, map->size, &dir)) rc = GNNSEG_EFORMAT;
    bool seen[6] = {false, false, false, false, false, false};
    int64_t len[4] = {0, 0, 0, 0};                                       // Ri_rows, Ri_cols, Ro_rows, Ro_cols, whatever their order in the archive
    static const char* names[6] = {"X.npy", "Ri_rows.npy", "Ri_cols.npy", "Ro_rows.npy", "Ro_cols.npy", "y.npy"};
    for (size_t i = 0; i < dir.size() && rc == GNNSEG_OK; ++i) {
        int which = -1;
        for (int k = 0; k < 6; ++k)
            if (dir[i].name == names[k]) which = k;
        if (which < 0) continue;                                         // other members are ignored
        if (dir[i].method != 0 || dir[i].csize != dir[i].usize) { rc = GNNSEG_EUNSUPPORTED; break; }   // np.savez_compressed
        // local file header: 30 bytes + name + extra, then the stored bytes
        const uint64_t lh = dir[i].offset;
        if (lh + 30 > map->size || rd32(b + lh) != 0x04034b50u) { rc = GNNSEG_EFORMAT; break; }
        const uint64_t data = lh + 30 + rd16(b + lh + 26) + rd16(b + lh + 28);
        if (data + dir[i].usize > map->size) { rc = GNNSEG_EFORMAT; break; }
        Member m;
        if (!parse_npy(b + data, dir[i].usize, &m)) { rc = GNNSEG_EFORMAT; break; }
        const bool is_x = which == 0, is_y = which == 5;
        // X: float32 (N, F); y: float32 (E,) (any width the reference may have stored is rejected
        // rather than converted: the packer consumes these pointers as they are); indices: int64 (E,)
        if (is_x ? !(m.kind == 'f' && m.item == 4 && m.ndim == 2)
                 : is_y ? !(m.kind == 'f' && m.item == 4 && m.ndim == 1)
                        : !(m.kind == 'i' && m.item == 8 && m.ndim == 1)) { rc = GNNSEG_EUNSUPPORTED; break; }
        const void* ptr = m.data;
        if (reinterpret_cast<uintptr_t>(ptr) % (uintptr_t)m.item != 0 && m.bytes > 0) {   // re-align once
            void* copy = nullptr;
            if (posix_memalign(&copy, 64, m.bytes) != 0) { rc = GNNSEG_EIO; break; }
            std::memcpy(copy, m.data, m.bytes);
            map->copies.push_back(copy);
            ptr = copy;
        }
        seen[which] = true;
        switch (which) {
            case 0: g->X = static_cast<const float*>(ptr); g->n_nodes = m.shape[0]; g->n_features = (int32_t)m.shape[1]; break;
            case 1: g->Ri_rows = static_cast<const int64_t*>(ptr); len[0] = m.shape[0]; break;
            case 2: g->Ri_cols = static_cast<const int64_t*>(ptr); len[1] = m.shape[0]; break;
            case 3: g->Ro_rows = static_cast<const int64_t*>(ptr); len[2] = m.shape[0]; break;
            case 4: g->Ro_cols = static_cast<const int64_t*>(ptr); len[3] = m.shape[0]; break;
            case 5: g->y = static_cast<const float*>(ptr); g->n_y = m.shape[0]; break;
        }
    }
    if (rc == GNNSEG_OK)
        for (int k = 0; k < 5; ++k)                                      // y may be absent; the five others may not
            if (!seen[k]) rc = GNNSEG_EFORMAT;
    if (rc == GNNSEG_OK && (len[0] != len[1] || len[2] != len[3])) rc = GNNSEG_EFORMAT;   // rows / cols of one matrix differ in length
    g->n_in = len[0];
    g->n_out = len[2];
    if (rc != GNNSEG_OK) {
        gnnseg_npz_close_graph_host(g);
        return rc;
    }
    return GNNSEG_OK;
}

extern "C" int gnnseg_npz_close_graph_host(GnnsegNpzGraph* g) {
    if (!g) return GNNSEG_EINVAL;
    Mapping* map = static_cast<Mapping*>(g->handle);
    if (map) {
        for (void* c : map->copies) std::free(c);
        if (map->base) munmap(map->base, map->size);
        delete map;
    }
    std::memset(g, 0, sizeof(*g));
    return GNNSEG_OK;
}

// A batch of files at once, opened by an OpenMP team (page-cache reads and header parsing in
// parallel).  graphs[b] is filled for every b; on the first failure all of them are closed again
// and that file's error code is returned.
extern "C" int gnnseg_npz_open_batch_host(int B, const char* const* paths, GnnsegNpzGraph* graphs, int n_threads) {
    if (B < 0 || (B > 0 && (!paths || !graphs))) return GNNSEG_EINVAL;
    std::vector<int> rcs((size_t)B, GNNSEG_OK);
    int nt = n_threads > 0 ? n_threads : 8;
    if (nt > B) nt = B > 0 ? B : 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int b = 0; b < B; ++b) rcs[b] = gnnseg_npz_open_graph_host(paths[b], &graphs[b]);
    for (int b = 0; b < B; ++b)
        if (rcs[b] != GNNSEG_OK) {
            for (int k = 0; k < B; ++k) gnnseg_npz_close_graph_host(&graphs[k]);
            return rcs[b];
        }
    return GNNSEG_OK;
}

extern "C" int gnnseg_npz_close_batch_host(int B, GnnsegNpzGraph* graphs) {
    if (B < 0 || (B > 0 && !graphs)) return GNNSEG_EINVAL;
    for (int b = 0; b < B; ++b) gnnseg_npz_close_graph_host(&graphs[b]);
    return GNNSEG_OK;
}
