/*
 * ffi_client.c -- a C client of the reference's FFI surface, run against libdivans_b200.so on the GPU box by
 * tests/test_gpu_c_client.py.  It drives the library the way the reference's own test client does (reference:
 * c/example.c, c/custom_alloc.h, c/arg.h -- not copied: the GPU box has no reference tree, and this file checks more):
 *   - every allocation of the library goes through a CAllocator that is either a checking malloc wrapper or a bump
 *     arena that only reclaims LIFO (the reference's NO_MALLOC mode); each block carries a header (magic, opaque,
 *     size) that free() verifies, blocks are 32-byte aligned (c/divans/ffi.h:39) and poisoned before hand-out, so the
 *     library must zero what it relies on (ffi/alloc_util.rs:72-83);
 *   - the decompressor constructor is declared with TWO parameters, exactly like the reference's C header
 *     (c/divans/ffi.h:61), although the library takes three (ffi/mod.rs:213): the third is whatever the register holds;
 *   - options through divans_set_option (reference selectors, ffi/interface.rs:18-37), including rejected ones;
 *   - streaming with caller buffers of 1, 15 and 65536 bytes on both sides; offsets are cursors;
 *   - NULL state / NULL offset pointers are DIVANS_FAILURE (ffi/mod.rs:241-262);
 *   - round trip equality; after every state is freed: no live block, and (arena mode: freed blocks below the top are
 *     not reclaimed, exactly like the reference's 255 MB test arena) a high-water mark far below 255 MB.
 * Exit code 0 = all checks passed; every failure prints a line and exits non-zero.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t DivansResult;
enum { DIVANS_SUCCESS = 0, DIVANS_NEEDS_MORE_INPUT = 1, DIVANS_NEEDS_MORE_OUTPUT = 2, DIVANS_FAILURE = 3 };
struct CAllocator { void *(*alloc_func)(void *opaque, size_t n); void (*free_func)(void *opaque, void *p); void *opaque; };
struct DivansDecompressorState; struct DivansCompressorState;
/* the reference header's (2-parameter) declaration */
struct DivansDecompressorState *divans_new_decompressor_with_custom_alloc(struct CAllocator alloc, uint8_t skip_crc);
struct DivansDecompressorState *divans_new_decompressor(void);
DivansResult divans_decode(struct DivansDecompressorState *, const uint8_t *, size_t, size_t *, uint8_t *, size_t, size_t *);
void divans_free_decompressor(struct DivansDecompressorState *);
struct DivansCompressorState *divans_new_compressor_with_custom_alloc(struct CAllocator alloc);
DivansResult divans_set_option(struct DivansCompressorState *, uint8_t selector, uint32_t value);
DivansResult divans_encode(struct DivansCompressorState *, const uint8_t *, size_t, size_t *, uint8_t *, size_t, size_t *);
DivansResult divans_encode_flush(struct DivansCompressorState *, uint8_t *, size_t, size_t *);
void divans_free_compressor(struct DivansCompressorState *);
uint8_t *divans_compressor_malloc_u8(struct DivansCompressorState *, size_t);
void divans_compressor_free_u8(struct DivansCompressorState *, uint8_t *, size_t);
size_t *divans_decompressor_malloc_usize(struct DivansDecompressorState *, size_t);
void divans_decompressor_free_usize(struct DivansDecompressorState *, size_t *, size_t);

#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

/* ---- checking allocator ---- */
#define MAGIC 0xD1FA57A7u
struct hdr { uint32_t magic; uint32_t arena; void *opaque; size_t size; void *raw; };
static int tag_heap, tag_arena;
static size_t live_blocks, live_bytes, total_allocs;
static unsigned char *arena; static size_t arena_cap, arena_top, arena_high;
static void *chk_alloc(void *opaque, size_t n) {
    size_t need = n + sizeof(struct hdr) + 64;
    unsigned char *raw;
    if (opaque == &tag_arena) {
        CHECK(arena_top + need <= arena_cap, "arena exhausted: top %zu need %zu cap %zu", arena_top, need, arena_cap);
        raw = arena + arena_top; arena_top += need; if (arena_top > arena_high) arena_high = arena_top;
    } else { CHECK(opaque == &tag_heap, "alloc: foreign opaque"); raw = (unsigned char *)malloc(need); CHECK(raw, "malloc"); }
    unsigned char *user = (unsigned char *)(((uintptr_t)raw + sizeof(struct hdr) + 31) & ~(uintptr_t)31);
    struct hdr *h = (struct hdr *)(user - sizeof(struct hdr));
    h->magic = MAGIC; h->arena = opaque == &tag_arena; h->opaque = opaque; h->size = n; h->raw = raw;
    memset(user, 0xCD, n);   /* poison: the library must not rely on fresh memory being zero without zeroing it itself */
    live_blocks++; live_bytes += n; total_allocs++;
    return user;
}
static void chk_free(void *opaque, void *p) {
    if (!p) return;
    struct hdr *h = (struct hdr *)((unsigned char *)p - sizeof(struct hdr));
    CHECK(h->magic == MAGIC, "free: pointer %p was not returned by alloc", p);
    CHECK(h->opaque == opaque, "free: wrong opaque");
    CHECK(live_blocks > 0, "free: double free");
    h->magic = 0; live_blocks--; live_bytes -= h->size;
    if (h->arena) { /* LIFO reclaim only, like a bump arena */
        size_t end = (size_t)((unsigned char *)h->raw - arena) + h->size + sizeof(struct hdr) + 64;
        if (end == arena_top) arena_top = (size_t)((unsigned char *)h->raw - arena);
    } else free(h->raw);
}

struct vec { uint8_t *p; size_t n, cap; };
static void vpush(struct vec *v, const uint8_t *d, size_t n) {
    if (v->n + n > v->cap) { v->cap = (v->n + n) * 2 + 64; v->p = (uint8_t *)realloc(v->p, v->cap); CHECK(v->p, "realloc"); }
    memcpy(v->p + v->n, d, n); v->n += n;
}

static void compress(struct CAllocator a, const uint8_t *data, size_t len, size_t in_step, size_t out_step, int literal_only, struct vec *out) {
    struct DivansCompressorState *st = divans_new_compressor_with_custom_alloc(a);
    CHECK(st, "divans_new_compressor_with_custom_alloc returned NULL (no GPU?)");
    CHECK(divans_set_option(st, 2 /* WINDOW_SIZE */, 20) == DIVANS_SUCCESS, "set window");
    CHECK(divans_set_option(st, 4 /* DYNAMIC_CONTEXT_MIXING */, 2) == DIVANS_SUCCESS, "set mixing");
    CHECK(divans_set_option(st, 11 /* PRIOR_DEPTH */, 1) == DIVANS_SUCCESS, "set prior depth");
    CHECK(divans_set_option(st, 8 /* LITERAL_ADAPTATION_CM_HIGH */, 6) == DIVANS_SUCCESS, "set adaptation");
    CHECK(divans_set_option(st, 8, 99) == DIVANS_FAILURE, "palette index 99 must be rejected");
    CHECK(divans_set_option(st, 9 /* FORCE_STRIDE_VALUE */, 77) == DIVANS_FAILURE, "stride 77 must be rejected");
    CHECK(divans_set_option(st, 200, 1) == DIVANS_FAILURE, "unknown selector must be rejected");
    if (literal_only) CHECK(divans_set_option(st, 5 /* USE_BROTLI_COMMAND_SELECTION */, 0) == DIVANS_SUCCESS, "set literal-only");
    uint8_t *scratch = divans_compressor_malloc_u8(st, 100);   /* the malloc helpers use the same allocator */
    CHECK(scratch, "compressor_malloc_u8"); divans_compressor_free_u8(st, scratch, 100);
    uint8_t *buf = (uint8_t *)malloc(out_step);
    size_t off = 0, dummy = 0;
    CHECK(divans_encode(NULL, data, len, &off, buf, out_step, &dummy) == DIVANS_FAILURE, "NULL state");
    CHECK(divans_encode(st, data, len, NULL, buf, out_step, &dummy) == DIVANS_FAILURE, "NULL input offset");
    while (off < len) {
        size_t chunk = len - off < in_step ? len - off : in_step, roff = 0, woff = 0;
        DivansResult r = divans_encode(st, data + off, chunk, &roff, buf, out_step, &woff);
        CHECK(r != DIVANS_FAILURE, "divans_encode failed");
        CHECK(roff <= chunk && woff <= out_step, "cursor past the window");
        CHECK(divans_set_option(st, 2, 22) == DIVANS_FAILURE, "options after the first encode must be rejected");
        off += roff; vpush(out, buf, woff);
    }
    DivansResult r;
    do {
        size_t woff = 0;
        r = divans_encode_flush(st, buf, out_step, &woff);
        CHECK(r != DIVANS_FAILURE && r != DIVANS_NEEDS_MORE_INPUT, "divans_encode_flush: %d", (int)r);
        CHECK(woff <= out_step, "flush cursor past the window");
        vpush(out, buf, woff);
    } while (r != DIVANS_SUCCESS);
    free(buf);
    divans_free_compressor(st);
}

static void decompress(struct CAllocator a, const uint8_t *data, size_t len, size_t in_step, size_t out_step, struct vec *out) {
    struct DivansDecompressorState *st = divans_new_decompressor_with_custom_alloc(a, 0);   /* two arguments, like c/example.c:62 */
    CHECK(st, "divans_new_decompressor_with_custom_alloc returned NULL (no GPU?)");
    size_t *sz = divans_decompressor_malloc_usize(st, 3); CHECK(sz, "malloc_usize"); divans_decompressor_free_usize(st, sz, 3);
    uint8_t *buf = (uint8_t *)malloc(out_step);
    size_t off = 0, dummy = 0;
    CHECK(divans_decode(st, data, len, &dummy, buf, out_step, NULL) == DIVANS_FAILURE, "NULL output offset");
    DivansResult r;
    do {
        size_t chunk = len - off < in_step ? len - off : in_step, roff = 0, woff = 0;
        r = divans_decode(st, data + off, chunk, &roff, buf, out_step, &woff);
        CHECK(r != DIVANS_FAILURE, "divans_decode failed at input offset %zu", off);
        CHECK(!(r == DIVANS_NEEDS_MORE_INPUT && off + roff == len && chunk == 0), "decoder wants input past the end of the stream");
        CHECK(roff <= chunk && woff <= out_step, "cursor past the window");
        off += roff; vpush(out, buf, woff);
    } while (r != DIVANS_SUCCESS);
    CHECK(off == len, "decoder left %zu input bytes unread", len - off);
    free(buf);
    divans_free_decompressor(st);
}

int main(int argc, char **argv) {
    const int use_arena = argc > 1 && strcmp(argv[1], "arena") == 0;
    if (use_arena) { arena_cap = (size_t)255 << 20; arena = (unsigned char *)malloc(arena_cap); CHECK(arena, "arena"); }
    struct CAllocator a = {chk_alloc, chk_free, use_arena ? (void *)&tag_arena : (void *)&tag_heap};
    /* input: pseudo-text with repeats, 300 kB + a binary tail (deterministic) */
    size_t len = 300000; uint8_t *data = (uint8_t *)malloc(len);
    uint32_t x = 12345;
    const char *words[] = {"the ", "quick ", "brown ", "fox ", "jumps ", "over ", "a ", "lazy ", "dog. ", "\n", "Mary ", "lamb "};
    size_t n = 0;
    while (n < len - 4096) { x = x * 1664525u + 1013904223u; const char *w = words[(x >> 24) % 12]; size_t l = strlen(w); memcpy(data + n, w, l); n += l; }
    while (n < len) { x = x * 1664525u + 1013904223u; data[n++] = (uint8_t)(x >> 23); }
    const size_t steps[][2] = {{65536, 65536}, {15, 4096}, {4096, 15}, {1, 65536}, {65536, 1}};
    for (int k = 0; k < 5; k++) {
        const size_t use = k < 3 ? len : 20000;                     /* 1-byte buffers: a shorter stream, same code path */
        struct vec comp = {0, 0, 0}, back = {0, 0, 0};
        compress(a, data, use, steps[k][0], steps[k][1], k & 1, &comp);
        CHECK(comp.n > 24 && comp.p[0] == 0xff && comp.p[1] == 0xe5 && comp.p[2] == 0x8c && comp.p[3] == 0x9f, "magic");
        CHECK(memcmp(comp.p + comp.n - 4, "ans~", 4) == 0, "trailer");
        decompress(a, comp.p, comp.n, steps[k][1], steps[k][0], &back);
        CHECK(back.n == use && memcmp(back.p, data, use) == 0, "round trip mismatch (in_step %zu out_step %zu)", steps[k][0], steps[k][1]);
        CHECK(live_blocks == 0 && live_bytes == 0, "%zu blocks (%zu bytes) leaked through the CAllocator", live_blocks, live_bytes);
        if (use_arena) CHECK(arena_high < ((size_t)128 << 20), "arena high-water mark %zu: the library allocates too much through the CAllocator", arena_high);
        printf("ok in_step=%zu out_step=%zu raw=%zu divans=%zu allocs=%zu arena_high=%zu\n", steps[k][0], steps[k][1], use, comp.n, total_allocs, arena_high);
        free(comp.p); free(back.p);
    }
    /* default constructor still works next to custom-allocator states */
    struct DivansDecompressorState *d = divans_new_decompressor(); CHECK(d, "divans_new_decompressor"); divans_free_decompressor(d);
    CHECK(total_allocs >= 20, "the library made only %zu allocations through the CAllocator", total_allocs);
    printf("all ok (%s allocator, %zu allocations)\n", use_arena ? "arena" : "heap", total_allocs);
    return 0;
}
