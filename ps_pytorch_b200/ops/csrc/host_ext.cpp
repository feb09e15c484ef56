// CPU-side native pieces (no CUDA):
//  * shuffle_lz_compress / shuffle_lz_decompress — lossless gradient/weight codec for the gloo tier: byte-plane
//    shuffle (all byte-0s, then all byte-1s, ... — exponent bytes of neighbouring fp32 values repeat) followed by an
//    LZ77 coder in the LZ4 block format.  Plays the role Blosc(shuffle)+snappy plays in the reference
//    (/root/reference/src/compression.py:18-46), in-tree instead of an external dependency.
//  * assemble_batch — one-pass gather + reflect-pad-4 random crop + horizontal flip + normalise from a resident uint8
//    dataset straight into a (pinned) fp32 batch buffer, NCHW or NHWC, multi-threaded; the data-loader hot loop
//    (reference: per-sample torchvision transforms, util.py:37-49, + vendored loader data_loader_ops/my_data_loader.py).
#include <torch/extension.h>
#include <ATen/Parallel.h>
#include <cstring>
#include <vector>

namespace psbh {

// ---------------------------------------------------------------------------------------------------- LZ4 block codec
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash4(uint32_t v) { return (v * 2654435761u) >> 18; }   // 14-bit table

static size_t lz_bound(size_t n) { return n + n / 255 + 32; }

static size_t lz_compress(const uint8_t* src, size_t n, uint8_t* dst) {
    std::vector<int32_t> table(1 << 14, -1);
    const uint8_t* ip = src; const uint8_t* anchor = src; const uint8_t* const iend = src + n;
    const uint8_t* const mflimit = n > 12 ? iend - 12 : src;     // last match must start before this
    uint8_t* op = dst;
    auto emit = [&](const uint8_t* lit, size_t litlen, size_t mlen, size_t offset, bool last) {
        uint8_t* token = op++;
        size_t l = litlen;
        if (l >= 15) { *token = 15 << 4; l -= 15; while (l >= 255) { *op++ = 255; l -= 255; } *op++ = (uint8_t)l; }
        else *token = (uint8_t)(l << 4);
        memcpy(op, lit, litlen); op += litlen;
        if (last) return;
        *op++ = (uint8_t)(offset & 0xff); *op++ = (uint8_t)(offset >> 8);
        size_t m = mlen - 4;
        if (m >= 15) { *token |= 15; m -= 15; while (m >= 255) { *op++ = 255; m -= 255; } *op++ = (uint8_t)m; }
        else *token |= (uint8_t)m;
    };
    if (n > 12) {
        while (ip < mflimit) {
            const uint32_t h = hash4(rd32(ip));
            const int32_t cand = table[h];
            table[h] = (int32_t)(ip - src);
            if (cand >= 0 && (ip - src) - cand <= 65535 && rd32(src + cand) == rd32(ip)) {
                const uint8_t* m = src + cand;
                size_t len = 4;
                const uint8_t* const lim = iend - 5;            // keep 5 trailing literals (format rule)
                while (ip + len < lim && ip[len] == m[len]) ++len;
                emit(anchor, (size_t)(ip - anchor), len, (size_t)(ip - m), false);
                ip += len; anchor = ip;
            } else {
                ++ip;
            }
        }
    }
    emit(anchor, (size_t)(iend - anchor), 0, 0, true);
    return (size_t)(op - dst);
}

static void lz_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t dn) {
    const uint8_t* ip = src; const uint8_t* const iend = src + n;
    uint8_t* op = dst; uint8_t* const oend = dst + dn;
    while (ip < iend) {
        const uint8_t token = *ip++;
        size_t l = token >> 4;
        if (l == 15) { uint8_t b; do { TORCH_CHECK(ip < iend, "lz: truncated"); b = *ip++; l += b; } while (b == 255); }
        TORCH_CHECK(ip + l <= iend && op + l <= oend, "lz: literal overrun");
        memcpy(op, ip, l); ip += l; op += l;
        if (ip >= iend) break;
        TORCH_CHECK(ip + 2 <= iend, "lz: truncated offset");
        const size_t off = ip[0] | ((size_t)ip[1] << 8); ip += 2;
        size_t m = token & 15;
        if (m == 15) { uint8_t b; do { TORCH_CHECK(ip < iend, "lz: truncated"); b = *ip++; m += b; } while (b == 255); }
        m += 4;
        TORCH_CHECK(off > 0 && (size_t)(op - dst) >= off && op + m <= oend, "lz: bad match");
        const uint8_t* mp = op - off;
        for (size_t i = 0; i < m; ++i) op[i] = mp[i];            // overlapping copy is intended
        op += m;
    }
    TORCH_CHECK(op == oend, "lz: size mismatch");
}

// blob = [u64 nblocks][u64 block_raw][nblocks x u64 comp_size][payloads]; blocks compress in parallel
static constexpr int64_t kBlock = 1 << 20;

at::Tensor shuffle_lz_compress(at::Tensor raw, int64_t itemsize) {
    TORCH_CHECK(raw.scalar_type() == at::kByte && raw.is_contiguous() && !raw.is_cuda());
    const int64_t n = raw.numel();
    TORCH_CHECK(itemsize >= 1 && n % itemsize == 0);
    const int64_t cnt = n / itemsize;
    std::vector<uint8_t> sh((size_t)n);
    const uint8_t* in = raw.data_ptr<uint8_t>();
    at::parallel_for(0, cnt, 1 << 16, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i)
            for (int64_t p = 0; p < itemsize; ++p) sh[(size_t)(p * cnt + i)] = in[i * itemsize + p];
    });
    const int64_t nb = (n + kBlock - 1) / kBlock;
    std::vector<std::vector<uint8_t>> outs((size_t)nb);
    at::parallel_for(0, nb, 1, [&](int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k) {
            const size_t len = (size_t)std::min<int64_t>(kBlock, n - k * kBlock);
            outs[(size_t)k].resize(lz_bound(len));
            outs[(size_t)k].resize(lz_compress(sh.data() + k * kBlock, len, outs[(size_t)k].data()));
        }
    });
    int64_t total = 16 + 8 * nb;
    for (auto& o : outs) total += (int64_t)o.size();
    at::Tensor blob = at::empty({total}, at::kByte);
    uint8_t* bp = blob.data_ptr<uint8_t>();
    uint64_t hdr[2] = {(uint64_t)nb, (uint64_t)kBlock};
    memcpy(bp, hdr, 16);
    uint8_t* cur = bp + 16 + 8 * nb;
    for (int64_t k = 0; k < nb; ++k) {
        uint64_t cs = outs[(size_t)k].size();
        memcpy(bp + 16 + 8 * k, &cs, 8);
        memcpy(cur, outs[(size_t)k].data(), cs);
        cur += cs;
    }
    return blob;
}

void shuffle_lz_decompress(at::Tensor blob, at::Tensor out, int64_t itemsize) {
    TORCH_CHECK(blob.scalar_type() == at::kByte && blob.is_contiguous() && out.scalar_type() == at::kByte && out.is_contiguous());
    const uint8_t* bp = blob.data_ptr<uint8_t>();
    uint64_t hdr[2];
    memcpy(hdr, bp, 16);
    const int64_t nb = (int64_t)hdr[0], blk = (int64_t)hdr[1], n = out.numel();
    TORCH_CHECK(nb == (n + blk - 1) / blk, "codec: block count mismatch");
    std::vector<size_t> offs((size_t)nb + 1);
    offs[0] = 16 + 8 * (size_t)nb;
    for (int64_t k = 0; k < nb; ++k) { uint64_t cs; memcpy(&cs, bp + 16 + 8 * k, 8); offs[(size_t)k + 1] = offs[(size_t)k] + cs; }
    TORCH_CHECK((int64_t)offs[(size_t)nb] == blob.numel(), "codec: payload size mismatch");
    std::vector<uint8_t> sh((size_t)n);
    at::parallel_for(0, nb, 1, [&](int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k)
            lz_decompress(bp + offs[(size_t)k], offs[(size_t)k + 1] - offs[(size_t)k], sh.data() + k * blk, (size_t)std::min<int64_t>(blk, n - k * blk));
    });
    const int64_t cnt = n / itemsize;
    uint8_t* o = out.data_ptr<uint8_t>();
    at::parallel_for(0, cnt, 1 << 16, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i)
            for (int64_t p = 0; p < itemsize; ++p) o[i * itemsize + p] = sh[(size_t)(p * cnt + i)];
    });
}

// ---------------------------------------------------------------------------------------------------- batch assembly
static inline int reflect(int i, int n) { if (i < 0) return -i; if (i >= n) return 2 * n - 2 - i; return i; }

// images u8 [N,C,H,W]; idx int64 [B]; out fp32 [B,C,H,W] (any strides: NCHW or channels_last); shifts in [0,8], flips u8
void assemble_batch(at::Tensor images, at::Tensor idx, at::Tensor out, at::Tensor mean, at::Tensor inv_std,
                    c10::optional<at::Tensor> dx, c10::optional<at::Tensor> dy, c10::optional<at::Tensor> flip) {
    TORCH_CHECK(images.scalar_type() == at::kByte && images.is_contiguous() && images.dim() == 4 && out.scalar_type() == at::kFloat);
    const int64_t B = idx.numel(), C = images.size(1), H = images.size(2), W = images.size(3);
    TORCH_CHECK(out.size(0) == B && out.size(1) == C && out.size(2) == H && out.size(3) == W);
    const uint8_t* im = images.data_ptr<uint8_t>();
    const int64_t* ix = idx.data_ptr<int64_t>();
    float* o = out.data_ptr<float>();
    const int64_t sb = out.stride(0), sc = out.stride(1), shh = out.stride(2), sw = out.stride(3);
    const float* mu = mean.data_ptr<float>(); const float* is = inv_std.data_ptr<float>();
    const bool aug = dx.has_value();
    const int64_t* pdx = aug ? dx->data_ptr<int64_t>() : nullptr;
    const int64_t* pdy = aug ? dy->data_ptr<int64_t>() : nullptr;
    const bool* pfl = aug ? flip->data_ptr<bool>() : nullptr;
    at::parallel_for(0, B, 8, [&](int64_t b0, int64_t b1) {
        for (int64_t b = b0; b < b1; ++b) {
            const uint8_t* src = im + ix[b] * C * H * W;
            const int ox = aug ? (int)pdx[b] - 4 : 0, oy = aug ? (int)pdy[b] - 4 : 0;
            const bool fl = aug && pfl[b];
            for (int64_t c = 0; c < C; ++c) {
                const float m = mu[c], s = is[c] * (1.0f / 255.0f), mm = m * is[c];
                for (int64_t h = 0; h < H; ++h) {
                    const int sy = reflect((int)h + oy, (int)H);
                    const uint8_t* row = src + (c * H + sy) * W;
                    float* orow = o + b * sb + c * sc + h * shh;
                    for (int64_t w = 0; w < W; ++w) {
                        const int wx = fl ? (int)(W - 1 - w) : (int)w;
                        const int sx = reflect(wx + ox, (int)W);
                        orow[w * sw] = (float)row[sx] * s - mm;
                    }
                }
            }
        }
    });
}

}  // namespace psbh

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("shuffle_lz_compress", &psbh::shuffle_lz_compress, pybind11::call_guard<pybind11::gil_scoped_release>());
    m.def("shuffle_lz_decompress", &psbh::shuffle_lz_decompress, pybind11::call_guard<pybind11::gil_scoped_release>());
    m.def("assemble_batch", &psbh::assemble_batch, pybind11::call_guard<pybind11::gil_scoped_release>());
}
