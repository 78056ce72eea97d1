// Session: device buffers + the launch sequences of the hot path
//   prep_audio (audio.rs:34-56) -> mel padding (transcribe.rs:161-177) -> forward_encoder
//   (mod.rs:228-260) -> cross K/V (mod.rs:484-485, hoisted out of the step loop) -> decoder steps.
// Windows of one call are batched: encoder rows of all windows are packed back to back.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "session.h"

namespace wb {

Session::Session(Model* model, int64_t max_w, int64_t max_b, int64_t max_text_len, int kv) : m(model) {
    if (!m || !m->finalized) fail(WB_ERR_STATE, "session: model not finalized");
    const wb_dims& D = m->dims;
    WB_REQUIRE(max_w >= 1 && max_b >= 1 && max_w * max_b <= 4096, "session: bad max_windows / max_beams");
    WB_REQUIRE(max_text_len >= 2 && max_text_len <= D.n_text_ctx, "session: max_text_len must be in [2, n_text_ctx]");
    WB_REQUIRE(kv == WB_KV_F32 || kv == WB_KV_F16, "session: kv_dtype must be WB_KV_F32 or WB_KV_F16");
    WB_CUDA(cudaSetDevice(m->device));
    max_windows = (int)max_w;
    max_beams = (int)max_b;
    t_max = (int)max_text_len;
    kv_dtype = kv;
    Rmax = max_windows * max_beams;
    TmS = D.n_audio_ctx + 2;
    Tcap = (D.n_audio_ctx - 1) / 2 + 1;
    Mcap = (int64_t)max_windows * Tcap;
    const int d = D.n_audio_state, H = D.n_text_head, L = D.n_text_layer, V = D.n_vocab;
    WB_REQUIRE(max_beams <= DEC_KC - 1, "session: max_beams must be <= 7 (candidates kept per record by the persistent decoders)");
    kmax = DEC_KC;

    WB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : ev) WB_CUDA(cudaEventCreate(&e));
    d_lmwin.alloc(max_windows); d_g1.alloc(max_windows); d_g2.alloc(max_windows); d_awin.alloc(max_windows);
    d_win_row_off.alloc(max_windows); d_win_T.alloc(max_windows);
    max_slots.alloc(max_windows);
    mel_rows.alloc((size_t)max_windows * TmS * N_MELS);
    x.alloc(Mcap * d);
    xa.alloc(Mcap * d);
    if (kv == WB_KV_F16) ckv16.alloc((size_t)L * Mcap * 2 * d); else ckv.alloc((size_t)L * Mcap * 2 * d);
    use_tc = m->fp16_exact;   // fp16-exact weights (released checkpoints): tensor-core encoder on fp16 hi/lo planes; else fp32 CUDA cores
    if (use_tc) {
        const size_t nm = (size_t)max_windows * TmS * N_MELS, nh = (size_t)max_windows * TmS * d;
        mel_h.alloc(nm); mel_l.alloc(nm); h1_h.alloc(nh); h1_l.alloc(nh);
        xn_h.alloc(Mcap * d); xn_l.alloc(Mcap * d); qkv_h.alloc(Mcap * 3 * d); qkv_l.alloc(Mcap * 3 * d);
        att_h.alloc(Mcap * d); att_l.alloc(Mcap * d); hid_h.alloc(Mcap * 4 * d); hid_l.alloc(Mcap * 4 * d);
        xa_h.alloc(Mcap * d); xa_l.alloc(Mcap * d);
        enc_plans.resize((size_t)2 + 4 * D.n_audio_layer + L);
        for (auto& pl : enc_plans) pl.reset(new GemmF16Plan());
    } else {
        h1.alloc((size_t)max_windows * TmS * d);
        xn.alloc(Mcap * d); att.alloc(Mcap * d); qkv.alloc(Mcap * 3 * d); hid.alloc(Mcap * 4 * d);
    }
    if (kv == WB_KV_F16) { kc16.alloc((size_t)L * Rmax * t_max * d); vc16.alloc((size_t)L * Rmax * t_max * d); }
    else { kc.alloc((size_t)L * Rmax * t_max * d); vc.alloc((size_t)L * Rmax * t_max * d); }
    dx.alloc((size_t)Rmax * d);
    att_pl.alloc(4 * dec5_plane_uint4(d)); hid_pl.alloc(8 * dec5_plane_uint4(d));   // attention + LayerNorm output planes (hi, lo each), MLP hidden planes
    WB_CUDA(cudaMemsetAsync(att_pl.p, 0, 4 * dec5_plane_uint4(d) * sizeof(uint4), st));
    WB_CUDA(cudaMemsetAsync(hid_pl.p, 0, 8 * dec5_plane_uint4(d) * sizeof(uint4), st));
    dq.alloc((size_t)Rmax * d); dhid.alloc((size_t)Rmax * 4 * d);
    logits.alloc((size_t)Rmax * V);
    tokens.alloc((size_t)Rmax * t_max); lengths.alloc(Rmax); cur_tok.alloc(Rmax); finished.alloc(Rmax);
    row_window.alloc(Rmax); anc0.alloc((size_t)Rmax * t_max); anc1.alloc((size_t)Rmax * t_max); parent.alloc(Rmax);
    pos.alloc(1); n_unfinished.alloc(128);
    topk_id.alloc((size_t)Rmax * kmax); topk_lp.alloc((size_t)Rmax * kmax);
    is_special.alloc(V);
    {
        const char* e = getenv("WB200_DECODER");   // "3": force the grid-barrier FMA fallback (decoder3.cu) for A/B runs
        if (e && e[0] == '3') dec_version = 3;
        ckv_hm = true;                             // cross K/V head-major (encoder.cu ckv_relayout_kernel): what the persistent decoders stream
        ckv_tmp.alloc(Mcap * 2 * d);
        cudaDeviceProp prop;
        WB_CUDA(cudaGetDeviceProperties(&prop, m->device));
        n_sm = prop.multiProcessorCount;
        n_logit_ctas = 2 * prop.multiProcessorCount;
        // persistent decoder: per-layer pointer table, barrier words, larger split-KV partial buffers
        part_o.alloc((size_t)Rmax * H * 16 * 64); part_m.alloc((size_t)Rmax * H * 16); part_l.alloc((size_t)Rmax * H * 16);
        datt.alloc((size_t)Rmax * d); steps_done.alloc(128); d3_bar.alloc(4);
        WB_CUDA(cudaMemsetAsync(d3_bar.p, 0, 4 * sizeof(unsigned int), st));
        { const char* e6 = getenv("WB200_DEC6"); use_dec6 = !(e6 && e6[0] == '0'); force_dec6 = e6 && e6[0] == 'f'; }
        {
            const bool h16 = m->fp16_exact;
            auto wp = [&](const LinearW& w) -> const void* { return h16 ? (const void*)w.w16 : (const void*)w.w32; };
            std::vector<Dec3Layer> lay((size_t)L);
            for (int l = 0; l < L; ++l) {
                const DecBlockW& B = m->dec[(size_t)l];
                Dec3Layer& y = lay[(size_t)l];
                y.ln1_g = B.attn_ln.g; y.ln1_b = B.attn_ln.b; y.ln1_eps = B.attn_ln.eps;
                y.ln2_g = B.cross_ln.g; y.ln2_b = B.cross_ln.b; y.ln2_eps = B.cross_ln.eps;
                y.ln3_g = B.mlp_ln.g; y.ln3_b = B.mlp_ln.b; y.ln3_eps = B.mlp_ln.eps;
                y.Wqkv = wp(B.qkv); y.Wo = wp(B.out); y.Wcq = wp(B.cq); y.Wco = wp(B.cout); y.W1 = wp(B.mlp1); y.W2 = wp(B.mlp2);
                y.bqkv = B.qkv.b; y.bo = B.out.b; y.bcq = B.cq.b; y.bco = B.cout.b; y.b1 = B.mlp1.b; y.b2 = B.mlp2.b;
            }
            d3_layers.alloc((size_t)L);
            WB_CUDA(cudaMemcpy(d3_layers.p, lay.data(), lay.size() * sizeof(Dec3Layer), cudaMemcpyHostToDevice));
            if (h16) {   // decoder5.cu stage descriptors
                auto gemm = [&](Dec5Desc& q, const void* Wp, const float* bias, int N, int n_slabs, int stage, int emit, int src) {
                    q.kind = D5_KIND_GEMM; q.W = Wp; q.bias = bias; q.N = N; q.n_slabs = n_slabs; q.stage = stage; q.emit = emit; q.src = src; q.ks = d;
                };
                auto ln = [&](Dec5Desc& q, const LayerNormW& w, int stage) {
                    q.kind = D5_KIND_LN; q.g = w.g; q.b = w.b; q.eps = w.eps; q.stage = stage; q.ks = d;
                };
                // The d x d projections (out, cross query, cross out) have only d/16 feature tiles -- 48 of 148 CTAs busy for small.en, each
                // staging all K columns of every row.  When d/256 slabs x d/16 tiles still fit ONE round of the grid, they run as K slabs of
                // 256 columns (one 32-column chunk per warp): three times the CTAs, a third of the staging each; the partial sums go to
                // ypart and are folded, in a fixed order, by the consumer (the next LayerNorm stage / the cross-attention query load).
                const int psl = d / 256;
                const char* e_split = getenv("WB200_D5_SPLIT");
                const bool can_split = !(e_split && e_split[0] == '0') && d % 256 == 0 && psl >= 2 && psl <= 4 && (d / 16) * psl <= n_sm;
                // Two tables: the split one is used by launches whose cross attention is NOT split over keys (n_splits == 1, i.e. at
                // least one (row, head) unit per SM: the batched shapes the split is for); small batches keep the unsplit stages, whose
                // cross-out staging merges the key-split partials over all d columns (launch_v3 picks per launch).
                auto build = [&](bool split_dd, DevBuf<Dec5Desc>& dst) {
                    std::vector<Dec5Desc> ds((size_t)L * 16 + 16);
                    for (int l = 0; l < L; ++l) {
                        const DecBlockW& B = m->dec[(size_t)l];
                        Dec5Desc* q = ds.data() + (size_t)l * 16;
                        ln(q[0], B.attn_ln, l == 0 ? D5_ST_LN_EMB : D5_ST_LN_FOLD);
                        gemm(q[1], B.qkv.w16, B.qkv.b, 3 * d, 1, D5_ST_PLANES, D5_EM_QKV, 3);
                        q[2].kind = D5_KIND_ATTN;
                        gemm(q[3], B.out.w16, B.out.b, d, 1, D5_ST_PLANES, D5_EM_RESID, 1);
                        ln(q[4], B.cross_ln, D5_ST_LN_X);
                        gemm(q[5], B.cq.w16, B.cq.b, d, 1, D5_ST_PLANES, D5_EM_CQ, 3);
                        q[6].kind = D5_KIND_ATTN;
                        gemm(q[7], B.cout.w16, B.cout.b, d, 1, D5_ST_CROSS, D5_EM_RESID, 1);
                        ln(q[8], B.mlp_ln, D5_ST_LN_X);
                        gemm(q[9], B.mlp1.w16, B.mlp1.b, 4 * d, 1, D5_ST_PLANES, D5_EM_HID, 3);
                        gemm(q[10], B.mlp2.w16, B.mlp2.b, d, 4, D5_ST_PLANES, D5_EM_PART, 2);
                        // MLP2 K = 4d: 3 slabs of 4d/3 when that keeps the 8-warp K split (multiple of 256, <= 1280): d/16 tiles x 3 slabs
                        // = 144 items for small.en -> ONE round on 148 CTAs instead of 192 items in two
                        if ((4 * d) % 3 == 0 && (4 * d / 3) % 256 == 0 && 4 * d / 3 <= 1280) { q[10].n_slabs = 3; q[10].ks = 4 * d / 3; }
                        if (split_dd) {
                            for (int sl : {3, 5, 7}) { q[sl].n_slabs = psl; q[sl].ks = 256; q[sl].emit = D5_EM_PART; }
                            q[4].stage = D5_ST_LN_FOLD; q[4].n_fold = psl; q[4].ks = 256;   // folds the out projection, feeds the split cross query
                            q[8].stage = D5_ST_LN_FOLD; q[8].n_fold = psl;                  // folds the cross out projection
                        }
                        if (l > 0) q[0].n_fold = q[10].n_slabs;                             // folds MLP2 of the previous layer
                    }
                    ln(ds[(size_t)L * 16 + 11], m->dec_ln, D5_ST_LN_FOLD_NOPUB);
                    ds[(size_t)L * 16 + 11].n_fold = ds[(size_t)(L - 1) * 16 + 10].n_slabs;
                    gemm(ds[(size_t)L * 16 + 12], m->tok_emb16, nullptr, V, 1, D5_ST_PLANES, D5_EM_LOGITS, 3);
                    dst.alloc(ds.size());
                    WB_CUDA(cudaMemcpy(dst.p, ds.data(), ds.size() * sizeof(Dec5Desc), cudaMemcpyHostToDevice));
                };
                build(false, d5_desc);
                if (can_split) build(true, d5_desc_split);
            }
        }
        ypart.alloc((size_t)4 * Rmax * d);   // decoder5.cu: MLP2 K-slab partial sums
        lg_m.alloc((size_t)n_logit_ctas * Rmax); lg_s.alloc((size_t)n_logit_ctas * Rmax);
        lg_v.alloc((size_t)n_logit_ctas * Rmax * DEC_KC); lg_i.alloc((size_t)n_logit_ctas * Rmax * DEC_KC);
    }
    WB_CUDA(cudaMallocHost((void**)&h_int, sizeof(int) * (4 * (size_t)Rmax + 16 + (size_t)Rmax * kmax)));
    WB_CUDA(cudaMallocHost((void**)&h_float, sizeof(float) * (size_t)Rmax * kmax));
    WB_CUDA(cudaMemsetAsync(is_special.p, 0, V, st));
    WB_CUDA(cudaStreamSynchronize(st));
}

Session::~Session() {
    if (st) cudaStreamSynchronize(st);
    for (auto& e : prof_ev) cudaEventDestroy(e);
    if (h_int) cudaFreeHost(h_int);
    if (h_float) cudaFreeHost(h_float);
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    if (st) cudaStreamDestroy(st);
}

// ---- geometry helpers -------------------------------------------------------------------------
static void set_geometry(Session& s, const std::vector<int>& Tm) {
    s.n_windows = (int)Tm.size();
    s.win_Tm = Tm;
    s.win_T.resize(Tm.size());
    s.win_row_off.resize(Tm.size());
    s.M_tot = 0;
    s.max_T = 0;
    s.max_Tm = 0;
    for (size_t w = 0; w < Tm.size(); ++w) {
        s.win_T[w] = (Tm[w] - 1) / 2 + 1;       // Conv1d k3 p1 s2 (mod.rs:179-182)
        s.win_row_off[w] = s.M_tot;
        s.M_tot += s.win_T[w];
        s.max_T = std::max(s.max_T, s.win_T[w]);
        s.max_Tm = std::max(s.max_Tm, Tm[w]);
    }
    const int d = s.m->dims.n_audio_state;
    std::vector<GemmGroup> g1(Tm.size()), g2(Tm.size());
    std::vector<AttnWindow> aw(Tm.size());
    for (size_t w = 0; w < Tm.size(); ++w) {
        // conv1: output frame t reads mel buffer rows t, t+1, t+2 (frames t-1, t, t+1), writes h1 row t+1
        g1[w] = GemmGroup{(int64_t)w * s.TmS * N_MELS, ((int64_t)w * s.TmS + 1) * d, Tm[w]};
        // conv2 (stride 2): output t reads h1 buffer rows 2t, 2t+1, 2t+2 (frames 2t-1, 2t, 2t+1)
        g2[w] = GemmGroup{(int64_t)w * s.TmS * d, s.win_row_off[w] * d, s.win_T[w]};
        aw[w] = AttnWindow{s.win_row_off[w], s.win_T[w]};
    }
    WB_CUDA(cudaMemcpyAsync(s.d_g1.p, g1.data(), g1.size() * sizeof(GemmGroup), cudaMemcpyHostToDevice, s.st));
    WB_CUDA(cudaMemcpyAsync(s.d_g2.p, g2.data(), g2.size() * sizeof(GemmGroup), cudaMemcpyHostToDevice, s.st));
    WB_CUDA(cudaMemcpyAsync(s.d_awin.p, aw.data(), aw.size() * sizeof(AttnWindow), cudaMemcpyHostToDevice, s.st));
    WB_CUDA(cudaMemcpyAsync(s.d_win_row_off.p, s.win_row_off.data(), Tm.size() * sizeof(int64_t),
                            cudaMemcpyHostToDevice, s.st));
    WB_CUDA(cudaMemcpyAsync(s.d_win_T.p, s.win_T.data(), Tm.size() * sizeof(int), cudaMemcpyHostToDevice, s.st));
    WB_CUDA(cudaStreamSynchronize(s.st));   // the host vectors above are temporaries
}

void Session::encode_from_device_wave(const float* wave_dev, const int64_t* offsets, const int64_t* lens, int64_t n) {
    WB_REQUIRE(n >= 1 && n <= max_windows, "encode: n_windows out of range for this session");
    const wb_dims& D = m->dims;
    std::vector<LogMelWindow> lw((size_t)n);
    std::vector<int> Tm((size_t)n);
    win_F.resize((size_t)n);
    int max_frames = 0;
    for (int64_t w = 0; w < n; ++w) {
        WB_REQUIRE(lens[w] >= N_FFT, "prep_audio: waveform shorter than n_fft (audio.rs:292)");
        WB_REQUIRE(lens[w] < (int64_t)1 << 30, "prep_audio: waveform too long");
        const int F = (int)(lens[w] / HOP);                        // frames after dropping the last one
        const int keep = std::min(F, D.n_audio_ctx - MEL_PADDING);  // transcribe.rs:173
        win_F[(size_t)w] = F;
        Tm[(size_t)w] = keep + MEL_PADDING;
        lw[(size_t)w] = LogMelWindow{offsets[w], (int)lens[w], F, keep, (int)w, ((int64_t)w * TmS + 1) * N_MELS};
        max_frames = std::max(max_frames, F);
    }
    WB_CUDA(cudaMemcpyAsync(d_lmwin.p, lw.data(), lw.size() * sizeof(LogMelWindow), cudaMemcpyHostToDevice, st));
    set_geometry(*this, Tm);   // syncs, so `lw` may go out of scope
    WB_CUDA(cudaEventRecord(ev[0], st));
    WB_CUDA(cudaMemsetAsync(mel_rows.p, 0, (size_t)n * TmS * N_MELS * sizeof(float), st));   // halo + 10 zero frames
    launch_logmel(*m, wave_dev, d_lmwin.p, (int)n, max_frames, mel_rows.p, max_slots.p, (int)n, st);
    WB_CUDA(cudaEventRecord(ev[1], st));
    run_encoder();
    WB_CUDA(cudaEventRecord(ev[2], st));
}

void Session::encode_waveforms_host(const float* const* waves, const int64_t* lens, int64_t n) {
    WB_REQUIRE(n >= 1 && n <= max_windows, "encode: n_windows out of range for this session");
    std::vector<int64_t> offs((size_t)n);
    int64_t total = 0;
    for (int64_t w = 0; w < n; ++w) {
        WB_REQUIRE(waves[w] != nullptr && lens[w] >= 0, "encode: null waveform");
        offs[(size_t)w] = total;
        total += (lens[w] + 3) / 4 * 4;   // keep windows 16-byte aligned
    }
    wave.ensure((size_t)total);
    for (int64_t w = 0; w < n; ++w)
        WB_CUDA(cudaMemcpyAsync(wave.p + offs[(size_t)w], waves[w], (size_t)lens[w] * sizeof(float),
                                cudaMemcpyHostToDevice, st));
    encode_from_device_wave(wave.p, offs.data(), lens, n);
}

void Session::encode_mels_host(const float* mel, int64_t n, int64_t n_mels, int64_t n_ctx) {
    const wb_dims& D = m->dims;
    WB_REQUIRE(n_mels == D.n_mels, "Audio mel spectrum size must be n_mels (mod.rs:231-235)");
    WB_REQUIRE(n_ctx >= 1 && n_ctx <= D.n_audio_ctx, "Audio length cannot exceed n_audio_ctx (mod.rs:236-241)");
    WB_REQUIRE(n >= 1 && n <= max_windows, "encode: n_windows out of range for this session");
    std::vector<int> Tm((size_t)n, (int)n_ctx);
    set_geometry(*this, Tm);
    DevBuf<float> tmp;
    tmp.alloc((size_t)n * n_mels * n_ctx);
    WB_CUDA(cudaMemcpyAsync(tmp.p, mel, tmp.n * sizeof(float), cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemsetAsync(mel_rows.p, 0, (size_t)n * TmS * N_MELS * sizeof(float), st));
    for (int64_t w = 0; w < n; ++w)
        launch_chan_to_rows(tmp.p + w * n_mels * n_ctx, mel_rows.p + ((int64_t)w * TmS + 1) * N_MELS, (int)n_ctx, n_ctx, st);
    run_encoder();
    WB_CUDA(cudaStreamSynchronize(st));
}

void Session::load_encoder_output_host(const float* xa_host, int64_t n, int64_t T) {
    WB_REQUIRE(n >= 1 && n <= max_windows && T >= 1 && T <= Tcap, "forward_decoder: encoder output shape out of range");
    std::vector<int> Tm((size_t)n, (int)(2 * T - 1));   // any Tm with (Tm-1)/2+1 == T
    set_geometry(*this, Tm);
    const int d = m->dims.n_audio_state;
    WB_CUDA(cudaMemcpyAsync(xa.p, xa_host, (size_t)n * T * d * sizeof(float), cudaMemcpyHostToDevice, st));
    if (use_tc) launch_split_f16(xa.p, xa_h.p, xa_l.p, (int64_t)n * T * d, st);
    run_cross_kv();
    encoded = true;
}

// ---- encoder ------------------------------------------------------------------------------------
void Session::run_encoder() {
    if (use_tc) run_encoder_f16();
    else run_encoder_f32();
    run_cross_kv();
    encoded = true;
}

// Tensor-core encoder (fp16-exact weights): conv stems, attention and MLP GEMMs are tcgen05 GEMMs over fp16 hi/lo planes
// (gemm_f16.cu), attention is enc_attn_tc.cu; only the residual stream x and the encoder output stay fp32 rows.
void Session::run_encoder_f16() {
    const wb_dims& D = m->dims;
    const int d = D.n_audio_state;
    const int M = (int)M_tot;
    const float qk_scale = (float)std::pow((double)d / (double)D.n_audio_head, -0.25);   // mod.rs:503
    auto run = [&](size_t site, const GemmF16Params& p, int64_t gstride, int rows_per_group) {
        GemmF16Plan& pl = *enc_plans[site];
        if (!pl.matches(p.max_rows, p.groups ? p.n_groups : 1)) pl.build(p, gstride, rows_per_group);   // tensor maps once per geometry
        pl.launch(st);
    };
    // halo rows of the conv1 output must read as zero padding
    WB_CUDA(cudaMemsetAsync(h1_h.p, 0, (size_t)n_windows * TmS * d * sizeof(__half), st));
    WB_CUDA(cudaMemsetAsync(h1_l.p, 0, (size_t)n_windows * TmS * d * sizeof(__half), st));
    launch_split_f16(mel_rows.p, mel_h.p, mel_l.p, (int64_t)n_windows * TmS * N_MELS, st);
    GemmF16Params p;
    // conv1 + GELU (mod.rs:243): K = 3*80 over three consecutive token-major mel rows; a conv output row is a dot product with
    // ONE contiguous 240-vector, so the conv is a GEMM whose A rows overlap (lda = 80 < K)
    p.A_hi = mel_h.p; p.A_lo = mel_l.p; p.lda = N_MELS; p.B = m->conv1.w16; p.P_hi = h1_h.p; p.P_lo = h1_l.p; p.ldc = d; p.N = d; p.K = 3 * N_MELS;
    p.bias = m->conv1.b; p.act = ACT_GELU; p.groups = d_g1.p; p.n_groups = n_windows; p.max_rows = max_Tm;
    run(0, p, (int64_t)TmS * N_MELS, TmS - 2);
    // conv2 (stride 2) + GELU + transpose + positional embedding (mod.rs:244-252)
    p = GemmF16Params{};
    p.A_hi = h1_h.p; p.A_lo = h1_l.p; p.lda = 2 * d; p.B = m->conv2.w16; p.C = x.p; p.ldc = d; p.N = d; p.K = 3 * d;
    p.bias = m->conv2.b; p.act = ACT_GELU; p.pos = m->enc_pos; p.groups = d_g2.p; p.n_groups = n_windows; p.max_rows = max_T;
    run(1, p, (int64_t)TmS * d, Tcap);
    for (int l = 0; l < D.n_audio_layer; ++l) {
        const EncBlockW& B = m->enc[(size_t)l];
        const size_t site = (size_t)2 + 4 * l;
        // x = x + attn(attn_ln(x))   (mod.rs:300)
        launch_layernorm_f16(x.p, nullptr, xn_h.p, xn_l.p, B.attn_ln, M, d, m->ln_eps_outside, st);
        p = GemmF16Params{};
        p.A_hi = xn_h.p; p.A_lo = xn_l.p; p.lda = d; p.B = B.qkv.w16; p.P_hi = qkv_h.p; p.P_lo = qkv_l.p; p.ldc = 3 * d; p.N = 3 * d; p.K = d;
        p.bias = B.qkv.b; p.scale = qk_scale; p.scale_cols = 2 * d; p.max_rows = M;
        run(site, p, 0, M);
        launch_encoder_attention_tc(qkv_h.p, qkv_l.p, att_h.p, att_l.p, d_awin.p, n_windows, max_T, d, D.n_audio_head, st);
        p = GemmF16Params{};
        p.A_hi = att_h.p; p.A_lo = att_l.p; p.lda = d; p.B = B.out.w16; p.C = x.p; p.ldc = d; p.N = d; p.K = d;
        p.bias = B.out.b; p.residual = x.p; p.max_rows = M;
        run(site + 1, p, 0, M);
        // x = x + mlp(mlp_ln(x))     (mod.rs:301); the MLP1 epilogue (bias + GELU) emits the hidden layer as planes
        launch_layernorm_f16(x.p, nullptr, xn_h.p, xn_l.p, B.mlp_ln, M, d, m->ln_eps_outside, st);
        p = GemmF16Params{};
        p.A_hi = xn_h.p; p.A_lo = xn_l.p; p.lda = d; p.B = B.mlp1.w16; p.P_hi = hid_h.p; p.P_lo = hid_l.p; p.ldc = 4 * d; p.N = 4 * d; p.K = d;
        p.bias = B.mlp1.b; p.act = ACT_GELU; p.max_rows = M;
        run(site + 2, p, 0, M);
        p = GemmF16Params{};
        p.A_hi = hid_h.p; p.A_lo = hid_l.p; p.lda = 4 * d; p.B = B.mlp2.w16; p.C = x.p; p.ldc = d; p.N = d; p.K = 4 * d;
        p.bias = B.mlp2.b; p.residual = x.p; p.max_rows = M;
        run(site + 3, p, 0, M);
    }
    launch_layernorm_f16(x.p, xa.p, xa_h.p, xa_l.p, m->ln_post, M, d, m->ln_eps_outside, st);   // mod.rs:259
}

// fp32 CUDA-core encoder: weights that are not exactly representable in fp16
void Session::run_encoder_f32() {
    const wb_dims& D = m->dims;
    const int d = D.n_audio_state;
    const int M = (int)M_tot;
    const float qk_scale = (float)std::pow((double)d / (double)D.n_audio_head, -0.25);   // mod.rs:503
    WB_CUDA(cudaMemsetAsync(h1.p, 0, (size_t)n_windows * TmS * d * sizeof(float), st));
    GemmParams p;
    p.A = mel_rows.p; p.lda = N_MELS; p.B = m->conv1.w32; p.C = h1.p; p.ldc = d; p.N = d; p.K = 3 * N_MELS;
    p.bias = m->conv1.b; p.act = ACT_GELU; p.groups = d_g1.p; p.n_groups = n_windows; p.max_rows = max_Tm;
    launch_gemm(p, st);
    p = GemmParams{};
    p.A = h1.p; p.lda = 2 * d; p.B = m->conv2.w32; p.C = x.p; p.ldc = d; p.N = d; p.K = 3 * d;
    p.bias = m->conv2.b; p.act = ACT_GELU; p.pos = m->enc_pos; p.groups = d_g2.p; p.n_groups = n_windows;
    p.max_rows = max_T;
    launch_gemm(p, st);
    for (int l = 0; l < D.n_audio_layer; ++l) {
        const EncBlockW& B = m->enc[(size_t)l];
        launch_layernorm(x.p, xn.p, B.attn_ln, M, d, m->ln_eps_outside, st);
        p = GemmParams{};
        p.A = xn.p; p.lda = d; p.B = B.qkv.w32; p.C = qkv.p; p.ldc = 3 * d; p.N = 3 * d; p.K = d;
        p.bias = B.qkv.b; p.scale = qk_scale; p.scale_cols = 2 * d; p.max_rows = M;
        launch_gemm(p, st);
        launch_encoder_attention(qkv.p, att.p, d_awin.p, n_windows, max_T, d, D.n_audio_head, st);
        p = GemmParams{};
        p.A = att.p; p.lda = d; p.B = B.out.w32; p.C = x.p; p.ldc = d; p.N = d; p.K = d;
        p.bias = B.out.b; p.residual = x.p; p.max_rows = M;
        launch_gemm(p, st);
        launch_layernorm(x.p, xn.p, B.mlp_ln, M, d, m->ln_eps_outside, st);
        p = GemmParams{};
        p.A = xn.p; p.lda = d; p.B = B.mlp1.w32; p.C = hid.p; p.ldc = 4 * d; p.N = 4 * d; p.K = d;
        p.bias = B.mlp1.b; p.act = ACT_GELU; p.max_rows = M;
        launch_gemm(p, st);
        p = GemmParams{};
        p.A = hid.p; p.lda = 4 * d; p.B = B.mlp2.w32; p.C = x.p; p.ldc = d; p.N = d; p.K = 4 * d;
        p.bias = B.mlp2.b; p.residual = x.p; p.max_rows = M;
        launch_gemm(p, st);
    }
    launch_layernorm(x.p, xa.p, m->ln_post, M, d, m->ln_eps_outside, st);   // mod.rs:259
}

// cross keys (pre-scaled) | values of every decoder layer, projected once per window (mod.rs:484-485 hoisted out of the step loop)
void Session::run_cross_kv() {
    const wb_dims& D = m->dims;
    const int d = D.n_text_state;
    const float qk_scale = (float)std::pow((double)d / (double)D.n_text_head, -0.25);
    for (int l = 0; l < D.n_text_layer; ++l) {
        const DecBlockW& B = m->dec[(size_t)l];
        float* c32 = ckv_hm ? ckv_tmp.p : (kv_dtype == WB_KV_F16 ? nullptr : ckv.p + (size_t)l * Mcap * 2 * d);
        if (use_tc) {
            GemmF16Params p;
            p.A_hi = xa_h.p; p.A_lo = xa_l.p; p.lda = d; p.B = B.ckv.w16; p.C = c32; p.ldc = 2 * d;
            p.N = 2 * d; p.K = d; p.bias = B.ckv.b; p.scale = qk_scale; p.scale_cols = d; p.max_rows = (int)M_tot;
            GemmF16Plan& pl = *enc_plans[(size_t)2 + 4 * D.n_audio_layer + l];
            if (!pl.matches(p.max_rows, 1)) pl.build(p, 0, (int)M_tot);
            pl.launch(st);
        } else {
            GemmParams p;
            p.A = xa.p; p.lda = d; p.B = B.ckv.w32; p.ldc = 2 * d;
            if (c32) p.C = c32; else p.C16 = ckv16.p + (size_t)l * Mcap * 2 * d;
            p.N = 2 * d; p.K = d; p.bias = B.ckv.b; p.scale = qk_scale; p.scale_cols = d; p.max_rows = (int)M_tot;
            launch_gemm(p, st);
        }
        if (ckv_hm) {
            void* dst = kv_dtype == WB_KV_F16 ? (void*)(ckv16.p + (size_t)l * Mcap * 2 * d) : (void*)(ckv.p + (size_t)l * Mcap * 2 * d);
            launch_ckv_relayout(ckv_tmp.p, dst, kv_dtype == WB_KV_F16, d_win_row_off.p, d_win_T.p, n_windows, M_tot, d, st);
        }
    }
}

// ---- decoder --------------------------------------------------------------------------------------
void Session::set_special(const uint8_t* sp) {
    if (sp) {
        WB_CUDA(cudaMemcpyAsync(is_special.p, sp, (size_t)m->dims.n_vocab, cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaStreamSynchronize(st));
        have_special = true;
    }
}

void Session::begin(const int64_t* prompt, int64_t prompt_len, bool prefill) {
    if (!encoded) fail(WB_ERR_STATE, "session: begin before encode");
    WB_REQUIRE(prompt_len >= 1 && prompt_len < t_max, "begin: prompt length out of range");
    const int V = m->dims.n_vocab;
    R = n_windows;
    WB_REQUIRE(R <= Rmax, "begin: too many rows");
    std::vector<int> tk((size_t)R * t_max, 0), rw((size_t)R), first((size_t)R);
    for (int r = 0; r < R; ++r) {
        for (int64_t i = 0; i < prompt_len; ++i) {
            WB_REQUIRE(prompt[i] >= 0 && prompt[i] < V, "begin: prompt token out of range");
            tk[(size_t)r * t_max + i] = (int)prompt[i];
        }
        rw[(size_t)r] = r;
        first[(size_t)r] = (int)prompt[0];
    }
    WB_CUDA(cudaMemcpyAsync(tokens.p, tk.data(), tk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemcpyAsync(row_window.p, rw.data(), rw.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemcpyAsync(cur_tok.p, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemsetAsync(finished.p, 0, sizeof(int) * Rmax, st));
    WB_CUDA(cudaMemsetAsync(pos.p, 0, sizeof(int), st));
    std::vector<int> len((size_t)R, (int)prompt_len);
    WB_CUDA(cudaMemcpyAsync(lengths.p, len.data(), len.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    anc_identity = true;
    anc_cur = 0;
    host_pos = 0;
    // feed prompt[0 .. prompt_len-1): no logits needed
    for (int64_t i = 0; prefill && i + 1 < prompt_len; ++i) {
        step_core(false, 0, 1, false, -1);
        std::vector<int> nxt((size_t)R, (int)prompt[i + 1]);
        WB_CUDA(cudaMemcpyAsync(cur_tok.p, nxt.data(), nxt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaStreamSynchronize(st));
    }
    WB_CUDA(cudaStreamSynchronize(st));
}

// One cooperative launch of the persistent decoder: n_steps positions starting at pos0.
void Session::launch_v3(int R_, int pos0, int n_steps, int logits_from, bool use_cur_tok, int mask_mode, int k,
                        bool greedy, int eot) {
    const wb_dims& D = m->dims;
    const int d = D.n_text_state, H = D.n_text_head;
    Dec3Args a;
    a.R = R_; a.Rmax = Rmax; a.d = d; a.H = H; a.L = D.n_text_layer; a.V = D.n_vocab; a.t_max = t_max; a.Mcap = Mcap;
    a.eps_outside = m->ln_eps_outside; a.qk_scale = (float)std::pow((double)d / (double)H, -0.25);
    a.layers = d3_layers.p; a.tok_emb = m->tok_emb32; a.pos_emb = m->dec_pos;
    a.E = m->fp16_exact ? (const void*)m->tok_emb16 : (const void*)m->tok_emb32;
    a.E_tiled = m->tok_emb16_tiled;
    a.lnf_g = m->dec_ln.g; a.lnf_b = m->dec_ln.b; a.lnf_eps = m->dec_ln.eps;
    a.x = dx.p; a.q = dq.p; a.att = datt.p; a.hid = dhid.p;
    a.ypart = ypart.p; a.lgbuf = logits.p; a.att_pl = att_pl.p; a.hid_pl = hid_pl.p; a.d5 = d5_desc.p; a.ckv_hm = ckv_hm ? 1 : 0;   // a.d5: see pick_d5 below (needs n_splits)
    a.lg_slices = std::max(1, std::min(16, n_sm / std::max(1, R_)));
    a.kv_half = kv_dtype == WB_KV_F16 ? 1 : 0;
    if (a.kv_half) { a.kc = kc16.p; a.vc = vc16.p; a.ckv = ckv16.p; } else { a.kc = kc.p; a.vc = vc.p; a.ckv = ckv.p; }
    a.row_window = row_window.p; a.win_row_off = d_win_row_off.p; a.win_T = d_win_T.p;
    a.anc = anc_identity ? nullptr : (anc_cur == 0 ? anc0.p : anc1.p);
    a.n_splits = std::max(1, std::min(16, n_sm / std::max(1, R_ * H)));
    auto pick_d5 = [&](int n_splits) { return (n_splits == 1 && d5_desc_split.p != nullptr) ? d5_desc_split.p : d5_desc.p; };
    a.d5 = pick_d5(a.n_splits);
    a.part_o = part_o.p; a.part_m = part_m.p; a.part_l = part_l.p;
    a.tokens = tokens.p; a.cur_tok = cur_tok.p; a.use_cur_tok = use_cur_tok ? 1 : 0;
    a.pos0 = pos0; a.n_steps = n_steps; a.logits_from = logits_from;
    a.is_special = have_special ? is_special.p : nullptr; a.mask_mode = mask_mode;
    a.k = k; a.greedy = greedy ? 1 : 0; a.eot = eot; a.lengths = lengths.p; a.finished = finished.p;
    a.topk_id = topk_id.p; a.topk_lp = topk_lp.p; a.logits_out = full_logits ? logits.p : nullptr;
    a.lg_m = lg_m.p; a.lg_s = lg_s.p; a.lg_v = lg_v.p; a.lg_i = lg_i.p;
    a.pos = pos.p; a.n_unfinished = n_unfinished.p; a.steps_done = steps_done.p; a.bar = d3_bar.p;
    if (getenv("WB200_TRACE")) {
        d3_trace.ensure(1 << 16);
        WB_CUDA(cudaMemsetAsync(d3_trace.p, 0, sizeof(unsigned long long) * (1 << 16), st));
        a.trace = d3_trace.p;
        a.trace_cap = 1 << 16;
    }
    WB_CUDA(cudaMemsetAsync(d3_bar.p, 0, 4 * sizeof(unsigned int), st));   // monotonic barrier counters start at 0
    last_decoder = 3;
    last_groups = 1;
    if (dec_version == 4) {
        // <= 7 rows: the cluster/DSMEM decoder (decoder4.cu, 131 us per position at 3 rows); 8..24 rows (batched chunks of the small
        // models): the head-fused tcgen05 cluster decoder (decoder6.cu), whose packed weight slices are built on first use.
        // WB200_DEC6=force sends the small batches through decoder6.cu too (tests, A/B runs).
        if (!force_dec6 && launch_dec4(a, m->fp16_exact, st)) last_decoder = 4;
        if (last_decoder == 3 && use_dec6 && m->fp16_exact && greedy && k == 1 && !use_cur_tok && a.anc == nullptr && a.logits_out == nullptr && R_ <= 24 && ckv_hm &&
            t_max <= 128 && dec6_supported(d, H)) {
            const int hs = dec6_pick_hs(d, R_);
            if (d6_pack[hs].p == nullptr) {
                std::vector<Dec6LayerSrc> src((size_t)D.n_text_layer);
                for (int l = 0; l < D.n_text_layer; ++l) {
                    const DecBlockW& B = m->dec[(size_t)l];
                    src[(size_t)l] = Dec6LayerSrc{B.qkv.w16, B.out.w16, B.cq.w16, B.cout.w16, B.mlp1.w16, B.mlp2.w16, B.qkv.b, B.out.b, B.cq.b, B.cout.b,
                                                  B.mlp1.b, B.mlp2.b, B.attn_ln.g, B.attn_ln.b, B.cross_ln.g, B.cross_ln.b, B.mlp_ln.g, B.mlp_ln.b,
                                                  B.attn_ln.eps, B.cross_ln.eps, B.mlp_ln.eps};
                }
                dec6_build_pack(d, hs, src, d6_pack[hs], d6_params[hs], st);
            }
            a.d6_pack = d6_pack[hs].p;
            a.d6_params = d6_params[hs].p;
            if (launch_dec6(a, hs, m->fp16_exact, st)) last_decoder = 6;
        }
        if (last_decoder == 3) {
            if (launch_dec4(a, m->fp16_exact, st)) last_decoder = 4;
            else if (launch_dec5(a, n_sm, m->fp16_exact, st)) last_decoder = 5;
            else if (R_ > 32 && m->fp16_exact && d % 256 == 0 && d <= 1280 && k <= DEC_KC) {
                // more rows than one launch of the batched tensor-core decoder takes (beams of many windows, BASELINE configs[4]:
                // 48 windows x 5 beams per GPU): row groups of 32, one launch each on the session stream; rows are independent,
                // ancestry entries stay absolute cache rows (a.kv_row0 = first cache row of the group)
                bool ok = true;
                int gi = 0;
                for (int r0 = 0; r0 < R_ && ok; r0 += 32, ++gi) {
                    const int Rg = std::min(32, R_ - r0);
                    Dec3Args g = a;
                    g.R = Rg; g.kv_row0 = r0;
                    g.x += (int64_t)r0 * d; g.q += (int64_t)r0 * d; g.att += (int64_t)r0 * d; g.hid += (int64_t)r0 * 4 * d;
                    g.row_window += r0;
                    if (g.anc) g.anc += (int64_t)r0 * t_max;
                    g.tokens += (int64_t)r0 * t_max; g.cur_tok += r0; g.lengths += r0; g.finished += r0;
                    g.topk_id += (int64_t)r0 * k; g.topk_lp += (int64_t)r0 * k;
                    if (g.logits_out) { g.logits_out += (int64_t)r0 * D.n_vocab; g.lgbuf = g.logits_out; }
                    g.lg_slices = std::max(1, std::min(16, n_sm / std::max(1, Rg)));
                    g.n_splits = std::max(1, std::min(16, n_sm / std::max(1, Rg * H)));
                    g.d5 = pick_d5(g.n_splits);
                    g.steps_done = steps_done.p + std::min(gi, 127); g.n_unfinished = n_unfinished.p + std::min(gi, 127);
                    WB_CUDA(cudaMemsetAsync(d3_bar.p, 0, 4 * sizeof(unsigned int), st));
                    ok = launch_dec5(g, n_sm, true, st);
                }
                if (!ok) fail(WB_ERR_UNSUPPORTED, "decoder5 rejected a row group");
                last_decoder = 5;
                last_groups = gi;
            }
        }
    }
    if (last_decoder == 3) launch_dec3(a, n_sm, m->fp16_exact, st);
    if (a.trace) {
        std::vector<unsigned long long> h(1 << 16);
        WB_CUDA(cudaStreamSynchronize(st));
        WB_CUDA(cudaMemcpy(h.data(), d3_trace.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(getenv("WB200_TRACE"), "w");   // debugging aid: WB200_TRACE=<file> receives the stage time stamps of CTA 0
        if (f) {
            for (size_t i = 0; i < h.size(); ++i)
                if (h[i]) fprintf(f, "%llu\n", h[i]);   // decoder6.cu: first half = consumer stage stamps, second half = MMA-warp stamps (time << 2 | kind)
            fclose(f);
        }
    }
}

void Session::step_core(bool with_logits, int mask_mode, int k, bool greedy, int eot) {
    WB_REQUIRE(k >= 1 && k <= DEC_KC - 1, "step: k must be in [1, 7]");
    launch_v3(R, host_pos, 1, with_logits ? 0 : INT_MAX, true, mask_mode, k, greedy, eot);
    ++host_pos;
}

void Session::profile_decode(const int64_t* prompt, int64_t prompt_len, int n_steps, int64_t eot, float* logits_ms,
                             float* step_ms) {
    WB_REQUIRE(n_steps >= 1 && prompt_len + n_steps <= t_max, "profile: n_steps out of range");
    while (prof_ev.size() < 2) {
        cudaEvent_t e;
        WB_CUDA(cudaEventCreate(&e));
        prof_ev.push_back(e);
    }
    // the whole decode is ONE kernel: time the launch (prefill + n_steps greedy steps) and report per position
    begin(prompt, prompt_len, false);
    const int total = (int)prompt_len - 1 + n_steps;
    WB_CUDA(cudaEventRecord(prof_ev[0], st));
    launch_v3(R, 0, total, (int)prompt_len - 1, false, 2, 1, true, -1 /* never stop early */);
    WB_CUDA(cudaEventRecord(prof_ev[1], st));
    WB_CUDA(cudaStreamSynchronize(st));
    float t = 0.f;
    WB_CUDA(cudaEventElapsedTime(&t, prof_ev[0], prof_ev[1]));
    (void)eot;
    *logits_ms = t / (float)total;
    *step_ms = t / (float)total;
}

void Session::step_beams(int64_t n_rows, const int32_t* window_of_row, const int32_t* parent_row, const int64_t* token,
                         int apply_mask, int k, int64_t* topk_ids_out, float* topk_lp_out) {
    if (!encoded) fail(WB_ERR_STATE, "session: step before encode/begin");
    WB_REQUIRE(n_rows >= 1 && n_rows <= Rmax, "step: n_rows out of range");
    WB_REQUIRE(k >= 1 && k <= kmax, "step: k out of range");
    WB_REQUIRE(host_pos + 1 < t_max, "step: session max_text_len exceeded");
    const int V = m->dims.n_vocab;
    int* hp = h_int;                  // parent
    int* hw = h_int + Rmax;           // window
    int* ht = h_int + 2 * Rmax;       // token
    for (int64_t r = 0; r < n_rows; ++r) {
        WB_REQUIRE(parent_row[r] >= 0 && parent_row[r] < R, "step: parent_row out of range");
        WB_REQUIRE(window_of_row[r] >= 0 && window_of_row[r] < n_windows, "step: window_of_row out of range");
        WB_REQUIRE(token[r] >= 0 && token[r] < V, "step: token out of range");
        hp[r] = parent_row[r];
        hw[r] = window_of_row[r];
        ht[r] = (int)token[r];
    }
    WB_CUDA(cudaMemcpyAsync(parent.p, hp, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemcpyAsync(row_window.p, hw, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemcpyAsync(cur_tok.p, ht, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st));
    if (anc_identity) {   // rows so far (prompt) live in their own cache rows
        launch_dec_anc_identity(anc0.p, Rmax, t_max, st);
        anc_cur = 0;
        anc_identity = false;
    }
    {
        int* cur = anc_cur == 0 ? anc0.p : anc1.p;
        int* nxt = anc_cur == 0 ? anc1.p : anc0.p;
        R = (int)n_rows;
        launch_dec_reorder(cur, nxt, parent.p, pos.p, R, t_max, st);
        anc_cur ^= 1;
    }
    step_core(true, apply_mask ? 1 : 0, k, false, -1);
    int* hid_ = h_int + 4 * Rmax + 16;
    WB_CUDA(cudaMemcpyAsync(hid_, topk_id.p, sizeof(int) * n_rows * k, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaMemcpyAsync(h_float, topk_lp.p, sizeof(float) * n_rows * k, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_rows * k; ++i) {
        topk_ids_out[i] = hid_[i];
        topk_lp_out[i] = h_float[i];
    }
}

void Session::greedy_decode(const int64_t* prompt, int64_t prompt_len, int max_depth, int64_t eot,
                            std::vector<std::vector<int64_t>>& out) {
    WB_REQUIRE(prompt_len + max_depth <= t_max, "greedy: prompt + max_depth exceeds the session's max_text_len");
    // one launch: prompt prefill + every greedy step, early exit inside the kernel
    begin(prompt, prompt_len, /*prefill=*/false);
    const int n_steps = (int)prompt_len - 1 + max_depth;
    if (max_depth > 0) launch_v3(R, 0, n_steps, (int)prompt_len - 1, false, 2, 1, true, (int)eot);
    std::vector<int> tk((size_t)R * t_max), len((size_t)R);
    int sdv[128] = {0};
    WB_CUDA(cudaMemcpyAsync(tk.data(), tokens.p, tk.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaMemcpyAsync(len.data(), lengths.p, len.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (max_depth > 0) WB_CUDA(cudaMemcpyAsync(sdv, steps_done.p, sizeof(int) * std::min(last_groups, 128), cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    int sd = 0;
    for (int g = 0; g < std::min(last_groups, 128); ++g) sd = std::max(sd, sdv[g]);   // row groups stop on their own
    last_steps = max_depth > 0 ? sd - ((int)prompt_len - 1) : 0;
    host_pos = (int)prompt_len - 1 + (int)last_steps;
    out.assign((size_t)R, {});
    for (int r = 0; r < R; ++r)
        for (int i = 0; i < len[(size_t)r]; ++i) out[(size_t)r].push_back(tk[(size_t)r * t_max + i]);
}

}  // namespace wb
