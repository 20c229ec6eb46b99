// owner.hpp -- library-owned host results: the public struct first, then what frees it.
#pragma once
#include <vector>

#include "../../include/xcltk_b200.h"

struct xg_reads_owner {
    xg_reads r;
    std::vector<void *> bufs;
    void (*free_fn)(void *) = nullptr;
};

struct xg_ctx;
struct xg_coo_owner {
    xg_coo m;
    std::vector<void *> bufs;   // pinned; returned to ctx's pool (or cudaFreeHost when ctx is null)
    xg_ctx *ctx = nullptr;
};
