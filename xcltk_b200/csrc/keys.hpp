// keys.hpp -- lossless 64-bit keys for cell barcodes / UMIs / query names (host side).
//
// The reference compares these as Python str (dict / set membership:
// xcltk/rdr/fc/mcount.py:34-43,119-127; xcltk/baf/fc/mcount.py:109-127,223-231).  The device
// compares 64-bit keys, so the mapping str -> key must be injective: short strings over
// {A,C,G,T,N,-,0-9} are bit-packed (reversible), everything else is interned.
// Layout documented in include/xcltk_b200.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/xcltk_b200.h"

namespace xg {

inline int key_char_code(unsigned char c) {
    switch (c) {
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 3;
        case 'T': return 4;
        case 'N': return 5;
        case '-': return 6;
        default: return (c >= '0' && c <= '9') ? 7 : -1;
    }
}

// Returns true and the packed key when `s` fits the 63-bit packed form.
inline bool key_pack(const char *s, int64_t n, uint64_t *out) {
    uint64_t k = 0;
    int bits = 0;
    for (int64_t i = 0; i < n; i++) {
        int code = key_char_code((unsigned char)s[i]);
        if (code < 0) return false;
        if (code < 7) {
            if (bits + 3 > 63) return false;
            k |= (uint64_t)code << (63 - bits - 3);
            bits += 3;
        } else {
            if (bits + 7 > 63) return false;
            k |= (uint64_t)((7 << 4) | (s[i] - '0')) << (63 - bits - 7);
            bits += 7;
        }
    }
    *out = k;
    return true;
}

inline int64_t key_unpack(uint64_t k, char *buf, int64_t cap) {
    static const char tab[] = "?ACGTN-";
    int bits = 0;
    int64_t n = 0;
    while (bits + 3 <= 63) {
        int code = (int)((k >> (63 - bits - 3)) & 7);
        if (code == 0) break;
        if (code < 7) {
            if (n < cap) buf[n] = tab[code];
            n++;
            bits += 3;
        } else {
            if (bits + 7 > 63) return -1;
            int d = (int)((k >> (63 - bits - 7)) & 15);
            if (d > 9) return -1;
            if (n < cap) buf[n] = (char)('0' + d);
            n++;
            bits += 7;
        }
    }
    return n;
}

}  // namespace xg

struct xg_keyspace {
    static const int NSHARD = 256;
    struct Shard {
        std::mutex mu;
        std::unordered_map<std::string, uint64_t> map;
        std::vector<std::string> names;
    };
    Shard shards[NSHARD];

    uint64_t intern(const char *s, int64_t n) {
        std::string str(s, (size_t)n);
        size_t h = std::hash<std::string>()(str);
        int si = (int)(h % NSHARD);
        Shard &sh = shards[si];
        std::lock_guard<std::mutex> g(sh.mu);
        auto it = sh.map.find(str);
        if (it != sh.map.end()) return it->second;
        uint64_t id = (uint64_t)sh.names.size() * NSHARD + (uint64_t)si;
        uint64_t key = (1ULL << 63) | id;
        sh.names.push_back(str);
        sh.map.emplace(std::move(str), key);
        return key;
    }

    uint64_t encode(const char *s, int64_t n) {
        uint64_t k;
        if (xg::key_pack(s, n, &k)) return k;
        return intern(s, n);
    }

    int64_t decode(uint64_t key, char *buf, int64_t cap) {
        if (key == XG_KEY_NONE || key == XG_KEY_NOMATCH) return -1;
        if (!(key >> 63)) return xg::key_unpack(key, buf, cap);
        uint64_t id = key & ~(1ULL << 63);
        Shard &sh = shards[id % NSHARD];
        std::lock_guard<std::mutex> g(sh.mu);
        uint64_t li = id / NSHARD;
        if (li >= sh.names.size()) return -1;
        const std::string &s = sh.names[li];
        int64_t n = (int64_t)s.size();
        memcpy(buf, s.data(), (size_t)(n < cap ? n : cap));
        return n;
    }

    int64_t n_interned() {
        int64_t t = 0;
        for (auto &sh : shards) {
            std::lock_guard<std::mutex> g(sh.mu);
            t += (int64_t)sh.names.size();
        }
        return t;
    }
};
