// Model container: receives the reference's npy-tree tensors (src/model/load.rs:19-310,
// python/dump.py:130-213), validates them against WhisperConfig (src/model/mod.rs:16-39) and
// re-lays them for the kernels:
//   * Linear [d_in, d_out] (burn layout, dump.py:141-145)  ->  [d_out][d_in] (k contiguous)
//   * query | key | value fused into one [3d][d] matrix; key has no bias (mod.rs:402-404) -> zeros
//   * Conv1d [out, in, k] (load.rs:145-161) -> [out][k*in + c] so that a conv output row is a dot
//     product with 3 consecutive token-major input rows
//   * decoder weights are stored in fp16 when every value is fp16-representable (OpenAI's released
//     checkpoints are fp16, so this is lossless for them); otherwise fp32.
#include <cmath>
#include <cstring>

#include "wb_internal.h"

namespace wb {

thread_local int64_t g_launch_count = 0;
static thread_local std::string g_last_error;

void set_last_error(const std::string& m) { g_last_error = m; }
const std::string& last_error_string() { return g_last_error; }
void fail(int code, const std::string& m) { throw Error(code, m); }

Model::~Model() {
    for (void* p : allocs) cudaFree(p);
}

namespace {

using Tensor = std::pair<std::vector<int64_t>, std::vector<float>>;

const Tensor& need(const Model& m, const std::string& path, std::initializer_list<int64_t> shape) {
    auto it = m.host.find(path);
    if (it == m.host.end()) fail(WB_ERR_STATE, "missing tensor: " + path);
    const auto& sh = it->second.first;
    std::vector<int64_t> want(shape);
    if (sh != want) {
        std::string s = "bad shape for " + path + ": got [";
        for (auto v : sh) s += std::to_string(v) + ",";
        s += "] want [";
        for (auto v : want) s += std::to_string(v) + ",";
        fail(WB_ERR_INVALID_ARG, s + "]");
    }
    return it->second;
}

bool fp16_exact(const std::vector<float>& v) {
    for (float f : v) {
        if (__half2float(__float2half_rn(f)) != f) return false;
    }
    return true;
}

float* up32(Model& m, const std::vector<float>& v) {
    float* p = nullptr;
    WB_CUDA(cudaMalloc((void**)&p, std::max<size_t>(v.size(), 1) * sizeof(float)));
    m.allocs.push_back(p);
    if (!v.empty()) WB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return p;
}
__half* up16(Model& m, const std::vector<float>& v) {
    std::vector<__half> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]);
    __half* p = nullptr;
    WB_CUDA(cudaMalloc((void**)&p, std::max<size_t>(v.size(), 1) * sizeof(__half)));
    m.allocs.push_back(p);
    if (!v.empty()) WB_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return p;
}

// burn [d_in][d_out] -> rows [d_out][d_in], appended to dst
void append_transposed(std::vector<float>& dst, const std::vector<float>& w, int d_in, int d_out) {
    const size_t base = dst.size();
    dst.resize(base + (size_t)d_in * d_out);
    for (int i = 0; i < d_in; ++i)
        for (int o = 0; o < d_out; ++o) dst[base + (size_t)o * d_in + i] = w[(size_t)i * d_out + o];
}

struct LinSpec {
    std::vector<float> w;   // [N][K]
    std::vector<float> b;   // [N]
    int n = 0, k = 0;
};

void add_linear(const Model& m, LinSpec& s, const std::string& path, int d_in, int d_out, bool has_bias) {
    const Tensor& w = need(m, path + "/weight", {d_in, d_out});
    append_transposed(s.w, w.second, d_in, d_out);
    if (has_bias) {
        const Tensor& b = need(m, path + "/bias", {d_out});
        s.b.insert(s.b.end(), b.second.begin(), b.second.end());
    } else {
        s.b.insert(s.b.end(), (size_t)d_out, 0.0f);
    }
    s.k = d_in;
    s.n += d_out;
}

LinearW upload_linear(Model& m, const LinSpec& s, bool want32, bool want16) {
    LinearW L;
    L.n = s.n;
    L.k = s.k;
    if (want32) L.w32 = up32(m, s.w);
    if (want16) L.w16 = up16(m, s.w);
    L.b = up32(m, s.b);
    return L;
}

LayerNormW upload_ln(Model& m, const std::string& path, int n) {
    LayerNormW L;
    L.g = up32(m, need(m, path + "/weight", {n}).second);
    L.b = up32(m, need(m, path + "/bias", {n}).second);
    auto it = m.host.find(path + "/eps");
    if (it == m.host.end() || it->second.second.empty()) fail(WB_ERR_STATE, "missing tensor: " + path + "/eps");
    L.eps = it->second.second.back();   // load.rs:47-49: scalars are stored as [1.0, value] / 1-element
    return L;
}

}  // namespace

void model_set_tensor(Model& m, const char* path, const float* data, const int64_t* shape, int ndim) {
    if (m.finalized) fail(WB_ERR_STATE, "model already finalized");
    WB_REQUIRE(path && data && ndim >= 0 && ndim <= 4, "set_tensor: bad arguments");
    size_t n = 1;
    std::vector<int64_t> sh;
    for (int i = 0; i < ndim; ++i) {
        WB_REQUIRE(shape[i] >= 0, "set_tensor: negative dimension");
        n *= (size_t)shape[i];
        sh.push_back(shape[i]);
    }
    std::string key(path);
    // scalars (LayerNorm eps) may arrive as 0-d or 1-element tensors
    if (key.size() >= 4 && key.compare(key.size() - 4, 4, "/eps") == 0) sh.clear();
    m.host[key] = Tensor(sh, std::vector<float>(data, data + n));
}

void model_finalize(Model& m) {
    if (m.finalized) fail(WB_ERR_STATE, "model already finalized");
    const wb_dims& D = m.dims;
    WB_REQUIRE(D.n_audio_state == D.n_text_state, "Audio encoder state size must be equal to text decoder state size.");
    WB_REQUIRE(D.n_audio_state % D.n_audio_head == 0 && D.n_text_state % D.n_text_head == 0,
               "State size must be a multiple of head size");
    WB_REQUIRE(D.n_audio_state / D.n_audio_head == 64 && D.n_text_state / D.n_text_head == 64,
               "only head dimension 64 (all Whisper sizes) is supported");
    WB_REQUIRE(D.n_mels == N_MELS, "n_mels must be 80");
    WB_REQUIRE(D.n_audio_state % 16 == 0, "n_state must be a multiple of 16");
    WB_CUDA(cudaSetDevice(m.device));
    const int d = D.n_audio_state;

    bool exact = true;
    for (const auto& kv : m.host) {
        if (kv.first.size() >= 4 && kv.first.compare(kv.first.size() - 4, 4, "/eps") == 0) continue;
        if (!fp16_exact(kv.second.second)) { exact = false; break; }
    }
    m.fp16_exact = exact;

    // ---- frontend tables
    const FrontendTables& ft = frontend_tables();
    m.basis_t = up32(m, ft.basis_t);
    m.mel_filt = up32(m, ft.mel_filt);
    {
        std::vector<int> rng(2 * N_MELS);
        for (int i = 0; i < N_MELS; ++i) { rng[2 * i] = ft.mel_lo[i]; rng[2 * i + 1] = ft.mel_hi[i]; }
        WB_CUDA(cudaMalloc((void**)&m.mel_range, rng.size() * sizeof(int)));
        m.allocs.push_back(m.mel_range);
        WB_CUDA(cudaMemcpy(m.mel_range, rng.data(), rng.size() * sizeof(int), cudaMemcpyHostToDevice));
    }

    // ---- encoder
    auto conv = [&](const std::string& path, int c_in) {
        const Tensor& w = need(m, path + "/weight", {d, c_in, 3});
        const Tensor& b = need(m, path + "/bias", {d});
        LinSpec s;
        s.n = d;
        s.k = 3 * c_in;
        s.w.resize((size_t)d * 3 * c_in);
        for (int o = 0; o < d; ++o)
            for (int c = 0; c < c_in; ++c)
                for (int kk = 0; kk < 3; ++kk)
                    s.w[(size_t)o * 3 * c_in + (size_t)kk * c_in + c] = w.second[((size_t)o * c_in + c) * 3 + kk];
        s.b = b.second;
        return upload_linear(m, s, !exact, exact);
    };
    m.conv1 = conv("encoder/conv1", D.n_mels);
    m.conv2 = conv("encoder/conv2", d);
    m.enc_pos = up32(m, need(m, "encoder/positional_embedding", {D.n_audio_ctx, d}).second);
    m.enc.resize(D.n_audio_layer);
    for (int i = 0; i < D.n_audio_layer; ++i) {
        const std::string p = "encoder/block_" + std::to_string(i);
        EncBlockW& B = m.enc[i];
        B.attn_ln = upload_ln(m, p + "/attn_ln", d);
        B.mlp_ln = upload_ln(m, p + "/mlp_ln", d);
        LinSpec qkv;
        add_linear(m, qkv, p + "/attn/query", d, d, true);
        add_linear(m, qkv, p + "/attn/key", d, d, false);
        add_linear(m, qkv, p + "/attn/value", d, d, true);
        B.qkv = upload_linear(m, qkv, !exact, exact);   // fp16-exact weights: fp16 storage feeds the tensor-core GEMM (gemm_f16.cu), else fp32 + SIMT
        LinSpec o, m1, m2;
        add_linear(m, o, p + "/attn/out", d, d, true);
        add_linear(m, m1, p + "/mlp/mlp1", d, 4 * d, true);
        add_linear(m, m2, p + "/mlp/mlp2", 4 * d, d, true);
        B.out = upload_linear(m, o, !exact, exact);
        B.mlp1 = upload_linear(m, m1, !exact, exact);
        B.mlp2 = upload_linear(m, m2, !exact, exact);
    }
    m.ln_post = upload_ln(m, "encoder/ln_post", d);

    // ---- decoder
    const bool h16 = exact, h32 = !exact;
    const Tensor& emb = need(m, "decoder/token_embedding/weight", {D.n_vocab, d});
    m.tok_emb32 = up32(m, emb.second);
    if (h16) m.tok_emb16 = up16(m, emb.second);
    if (h16 && (d == 128 || d == 384)) {
        // decoder4.cu streams the logits matrix as contiguous half-tiles [tile of 16 rows][K half][16][d/2] (one bulk copy each,
        // rows >= V are zero) and feeds them to mma.sync from shared memory
        const int V = D.n_vocab, tiles = (V + 15) / 16, kh = d / 2;
        std::vector<float> tl((size_t)tiles * 2 * 16 * kh, 0.0f);
        for (int t = 0; t < tiles; ++t)
            for (int hh = 0; hh < 2; ++hh)
                for (int rr = 0; rr < 16 && t * 16 + rr < V; ++rr)
                    std::memcpy(&tl[(((size_t)t * 2 + hh) * 16 + rr) * kh], &emb.second[(size_t)(t * 16 + rr) * d + (size_t)hh * kh], sizeof(float) * kh);
        m.tok_emb16_tiled = up16(m, tl);
    }
    m.dec_pos = up32(m, need(m, "decoder/positional_embedding", {D.n_text_ctx, d}).second);
    m.dec.resize(D.n_text_layer);
    for (int i = 0; i < D.n_text_layer; ++i) {
        const std::string p = "decoder/block_" + std::to_string(i);
        DecBlockW& B = m.dec[i];
        B.attn_ln = upload_ln(m, p + "/attn_ln", d);
        B.cross_ln = upload_ln(m, p + "/cross_attn_ln", d);
        B.mlp_ln = upload_ln(m, p + "/mlp_ln", d);
        LinSpec qkv, o, cq, ckv, co, m1, m2;
        add_linear(m, qkv, p + "/attn/query", d, d, true);
        add_linear(m, qkv, p + "/attn/key", d, d, false);
        add_linear(m, qkv, p + "/attn/value", d, d, true);
        add_linear(m, o, p + "/attn/out", d, d, true);
        add_linear(m, cq, p + "/cross_attn/query", d, d, true);
        add_linear(m, ckv, p + "/cross_attn/key", d, d, false);
        add_linear(m, ckv, p + "/cross_attn/value", d, d, true);
        add_linear(m, co, p + "/cross_attn/out", d, d, true);
        add_linear(m, m1, p + "/mlp/mlp1", d, 4 * d, true);
        add_linear(m, m2, p + "/mlp/mlp2", 4 * d, d, true);
        B.qkv = upload_linear(m, qkv, h32, h16);
        B.out = upload_linear(m, o, h32, h16);
        B.cq = upload_linear(m, cq, h32, h16);
        B.ckv = upload_linear(m, ckv, h32, h16);      // applied by the encoder-side GEMM
        B.cout = upload_linear(m, co, h32, h16);
        B.mlp1 = upload_linear(m, m1, h32, h16);
        B.mlp2 = upload_linear(m, m2, h32, h16);
    }
    m.dec_ln = upload_ln(m, "decoder/ln", d);
    {
        std::vector<float> z((size_t)std::max(4 * d, D.n_vocab), 0.0f);
        m.zero_bias = up32(m, z);
    }
    m.host.clear();
    m.finalized = true;
}

}  // namespace wb
