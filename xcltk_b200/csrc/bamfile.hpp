// bamfile.hpp -- BGZF container + BAM header, shared by the host decoder (decode.cpp) and the
// device decoder (gpu_decode.cu).  Definitions live in decode.cpp.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace xg_dec {

struct BgzfBlock {
    uint64_t coff;    // offset of the deflate payload in the file
    uint32_t clen;    // deflate payload length
    uint32_t isize;   // uncompressed length
    uint64_t uoff;    // offset in the uncompressed stream
    uint32_t crc;     // CRC32 of the uncompressed bytes (gzip trailer)
};

// Byte buffer without the value-initialisation of std::vector::resize (GBs of memset).
struct Bytes {
    std::unique_ptr<uint8_t[]> p;
    size_t n = 0;
    void resize(size_t m) {
        p.reset(new uint8_t[m ? m : 1]);
        n = m;
    }
    size_t size() const { return n; }
    uint8_t *data() { return p.get(); }
    const uint8_t *data() const { return p.get(); }
    uint8_t &operator[](size_t i) { return p[i]; }
    const uint8_t &operator[](size_t i) const { return p[i]; }
};

struct Header {
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    uint64_t end_off = 0;   // offset of the first record in the uncompressed stream
};

// Raw file + block index + parsed header (only the blocks the header spans are inflated).
struct BamFile {
    Bytes f;
    std::vector<BgzfBlock> blocks;
    Header h;
};

// All return XG_OK or an XG_E_* code with the message left for xg_host_last_error().
int read_file(const char *path, Bytes &buf, int64_t max_bytes = -1);
int scan_bgzf(const Bytes &f, std::vector<BgzfBlock> &blocks, const char *path, bool allow_partial_tail = false);
int bgzf_block_header(const uint8_t *p, uint64_t avail, uint32_t *total, uint32_t *hdr_len);
int inflate_blocks(const uint8_t *f, uint64_t f_base, const BgzfBlock *blocks, size_t n_blocks, Bytes &out,
                   int n_threads);
int parse_header(const Bytes &u, Header &h, const char *path);
int open_bam(const char *path, BamFile &out, bool header_only);
int fail(int code, const std::string &msg);

}  // namespace xg_dec
