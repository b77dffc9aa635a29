/*
 * CPU oracle for the speculative-decoding verify hot path -- plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/libsd_oracle.so by
 * oracle/Makefile (and by __graft_entry__.build()).  Only tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * library (llm-inference-lab_b200/csrc) never links or calls it.
 *
 * Each function restates the algorithm of the reference repository
 * GogoRit/llm-inference-lab (paths relative to the reference root); parity with
 * the reference is pinned by tests/golden/ (see tests/test_oracle.py).
 *
 * dtype codes are shared with include/sd_verify.h:
 *   0 = fp32, 1 = fp16, 2 = bf16;   id dtype: 0 = int32, 1 = int64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float load_as_f32(const void* base, int dtype, int64_t i) {
    if (dtype == 0) return ((const float*)base)[i];
    if (dtype == 1) return (float)((const _Float16*)base)[i];
    uint32_t w = (uint32_t)((const uint16_t*)base)[i] << 16;
    float f;
    memcpy(&f, &w, 4);
    return f;
}

static inline int64_t load_id(const void* base, int id_dtype, int64_t i) {
    return id_dtype == 0 ? (int64_t)((const int32_t*)base)[i] : ((const int64_t*)base)[i];
}

/* torch.argmax semantics (reference: src/kernels/reference.py:37): first maximal
 * element; NaN is maximal and the first NaN wins; -0.0 == +0.0. */
static int64_t argmax_row(const void* row, int dtype, int64_t n) {
    int64_t best_i = 0;
    float best = load_as_f32(row, dtype, 0);
    if (isnan(best)) return 0;
    for (int64_t i = 1; i < n; ++i) {
        float v = load_as_f32(row, dtype, i);
        if (isnan(v)) return i;
        if (v > best) {
            best = v;
            best_i = i;
        }
    }
    return best_i;
}

int sd_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * verify_prefix (+ bonus when rows == K + 1).
 * reference: src/kernels/reference.py:13-56; bonus row selection:
 * src/specdec/core/pipeline.py:3140-3163, :3174-3206, :3341-3362.
 * logits element strides: stride_b, stride_r (innermost stride 1).
 * target_ids [B, rows] and bonus [B] may be NULL.
 */
int sd_oracle_verify(const void* logits, int dtype, int64_t stride_b, int64_t stride_r,
                     const void* ids, int id_dtype, int64_t id_stride_b, int B, int rows, int K,
                     int V, int32_t* accept_len, uint8_t* mask, int32_t* target_ids,
                     int32_t* bonus) {
    if (B < 0 || rows < K || rows > K + 1 || V <= 0) return 1;
    size_t esz = dtype == 0 ? 4 : 2;
    int64_t total = (int64_t)B * rows;
    int32_t* amax = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    if (!amax) return 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < total; ++r) {
        int64_t b = r / rows, j = r % rows;
        const char* row = (const char*)logits + (size_t)(b * stride_b + j * stride_r) * esz;
        amax[r] = (int32_t)argmax_row(row, dtype, V);
    }
    for (int b = 0; b < B; ++b) {
        int n = 0;
        while (n < K && (int64_t)amax[(int64_t)b * rows + n] == load_id(ids, id_dtype, b * id_stride_b + n)) ++n;
        accept_len[b] = n;
        for (int k = 0; k < K; ++k) mask[(int64_t)b * K + k] = (uint8_t)(k < n);
        if (target_ids)
            for (int j = 0; j < rows; ++j) target_ids[(int64_t)b * rows + j] = amax[(int64_t)b * rows + j];
        if (bonus) bonus[b] = rows == K + 1 ? amax[(int64_t)b * rows + n] : -1;
    }
    free(amax);
    return 0;
}

/*
 * kv_append_with_mask on raw bytes (row_bytes = D * element size).
 * reference: src/kernels/reference.py:96-159.  One call handles K or V.
 * base [B,H,L,row], draft [B,H,K,row] -> out [B,H,L+K,row].
 * mask == NULL means "plain concatenation" (reference.py:59-93).
 */
int sd_oracle_kv_append(const uint8_t* base, const uint8_t* draft, const uint8_t* mask,
                        const int32_t* accept_len, uint8_t* out, int B, int H, int L, int K,
                        int64_t row_bytes) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int h = 0; h < H; ++h) {
            const uint8_t* src = base + ((int64_t)b * H + h) * L * row_bytes;
            const uint8_t* dr = draft + ((int64_t)b * H + h) * K * row_bytes;
            uint8_t* dst = out + ((int64_t)b * H + h) * (L + K) * row_bytes;
            memcpy(dst, src, (size_t)(L * row_bytes));
            uint8_t* tail = dst + (int64_t)L * row_bytes;
            if (!mask) {
                memcpy(tail, dr, (size_t)(K * row_bytes));
                continue;
            }
            memset(tail, 0, (size_t)(K * row_bytes));
            int want = accept_len[b], got = 0;
            for (int k = 0; k < K && got < want; ++k) {
                if (mask[(int64_t)b * K + k]) {
                    memcpy(tail + (int64_t)got * row_bytes, dr + (int64_t)k * row_bytes, (size_t)row_bytes);
                    ++got;
                }
            }
        }
    }
    return 0;
}

/*
 * In-place arena append (the B200 layout): arena [B,H,C,row]; sequence b holds
 * seq_lens[b] rows; the first n_keep[b] of the R offered rows are written at
 * seq_lens[b].., the other offered slots are zeroed if zero_rejected.
 * Restates reference.py:137-157 for a prefix mask.  new_lens may be NULL.
 */
int sd_oracle_kv_arena_append(uint8_t* arena, const int32_t* seq_lens, const uint8_t* fresh,
                              const int32_t* n_keep, int32_t* new_lens, int B, int H, int C, int R,
                              int64_t row_bytes, int zero_rejected) {
    for (int b = 0; b < B; ++b)
        if (seq_lens[b] < 0 || seq_lens[b] + R > C) return 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int h = 0; h < H; ++h) {
            int n = n_keep[b] < 0 ? 0 : (n_keep[b] > R ? R : n_keep[b]);
            uint8_t* dst = arena + (((int64_t)b * H + h) * C + seq_lens[b]) * row_bytes;
            const uint8_t* src = fresh + ((int64_t)b * H + h) * R * row_bytes;
            memcpy(dst, src, (size_t)(n * row_bytes));
            if (zero_rejected) memset(dst + (int64_t)n * row_bytes, 0, (size_t)((R - n) * row_bytes));
        }
    }
    if (new_lens)
        for (int b = 0; b < B; ++b) {
            int n = n_keep[b] < 0 ? 0 : (n_keep[b] > R ? R : n_keep[b]);
            new_lens[b] = seq_lens[b] + n;
        }
    return 0;
}
