// Reader for the reference's model-file format: the npy tree written by python/dump.py:120-213 and read by
// src/model/load.rs:19-310.  Every tensor is ONE .npy file of dtype f32 whose payload is
// [dim_0, ..., dim_{D-1}, values...] (load.rs:19-27 numpy_to_tensor: the first D floats are the shape);
// scalars are stored as [1.0, value] (dump.py:130-132, load.rs:47-53).  Linear weights are burn layout
// [d_in, d_out] (dump.py:141-145).  The tree also carries the model dimensions (n_layer, n_mels,
// n_audio_state, n_head), so a model can be created from the directory alone like load::load_whisper does.
#include <cstdio>
#include <cstring>
#include <fstream>

#include "session.h"

namespace wb {

namespace {

std::vector<float> read_npy_f32(const std::string& file) {
    std::ifstream f(file, std::ios::binary);
    if (!f) fail(WB_ERR_STATE, "cannot open " + file);
    char magic[8];
    f.read(magic, 8);
    if (!f || std::memcmp(magic, "\x93NUMPY", 6) != 0) fail(WB_ERR_INVALID_ARG, "not an npy file: " + file);
    const int major = (unsigned char)magic[6];
    uint32_t hlen = 0;
    if (major == 1) {
        unsigned char b[2];
        f.read((char*)b, 2);
        hlen = b[0] | (b[1] << 8);
    } else {
        unsigned char b[4];
        f.read((char*)b, 4);
        hlen = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    }
    std::string hdr(hlen, '\0');
    f.read(&hdr[0], hlen);
    if (!f) fail(WB_ERR_INVALID_ARG, "truncated npy header: " + file);
    if (hdr.find("'<f4'") == std::string::npos && hdr.find("'|f4'") == std::string::npos && hdr.find("'=f4'") == std::string::npos)
        fail(WB_ERR_INVALID_ARG, "npy dtype must be float32 (npy::NpyData<f32>, load.rs:19): " + file);
    if (hdr.find("'fortran_order': True") != std::string::npos) fail(WB_ERR_INVALID_ARG, "fortran-order npy not supported: " + file);
    const size_t sp = hdr.find("'shape':");
    if (sp == std::string::npos) fail(WB_ERR_INVALID_ARG, "npy header without shape: " + file);
    const size_t lp = hdr.find('(', sp), rp = hdr.find(')', sp);
    size_t count = 1;
    {
        std::string dims = hdr.substr(lp + 1, rp - lp - 1);
        size_t pos = 0;
        bool any = false;
        while (pos < dims.size()) {
            while (pos < dims.size() && (dims[pos] == ' ' || dims[pos] == ',')) ++pos;
            if (pos >= dims.size()) break;
            count *= (size_t)std::strtoull(dims.c_str() + pos, nullptr, 10);
            any = true;
            while (pos < dims.size() && dims[pos] != ',') ++pos;
        }
        if (!any) count = 1;
    }
    std::vector<float> v(count);
    f.read((char*)v.data(), (std::streamsize)(count * sizeof(float)));
    if (!f) fail(WB_ERR_INVALID_ARG, "truncated npy payload: " + file);
    return v;
}

// load.rs:19-27: first `rank` floats are the dims
void read_tensor(const std::string& dir, const std::string& path, int rank, std::vector<int64_t>& shape, std::vector<float>& vals) {
    const std::vector<float> v = read_npy_f32(dir + "/" + path + ".npy");
    if ((int)v.size() < rank) fail(WB_ERR_INVALID_ARG, "npy tensor too short: " + path);
    shape.clear();
    size_t n = 1;
    for (int i = 0; i < rank; ++i) {
        shape.push_back((int64_t)v[(size_t)i]);
        n *= (size_t)v[(size_t)i];
    }
    if (v.size() != (size_t)rank + n) fail(WB_ERR_INVALID_ARG, "npy tensor size does not match its leading dims: " + path);
    vals.assign(v.begin() + rank, v.end());
}

float read_scalar(const std::string& dir, const std::string& path) {   // stored as [1.0, value]
    const std::vector<float> v = read_npy_f32(dir + "/" + path + ".npy");
    if (v.size() != 2) fail(WB_ERR_INVALID_ARG, "npy scalar must be [1.0, value]: " + path);
    return v[1];
}

}  // namespace

// dims from the tree alone (host only)
void npy_tree_probe(const std::string& dir, wb_dims& D) {
    std::vector<int64_t> sh;
    std::vector<float> vals;
    D.n_mels = (int)read_scalar(dir, "encoder/n_mels");
    D.n_audio_state = (int)read_scalar(dir, "encoder/n_audio_state");
    D.n_audio_layer = (int)read_scalar(dir, "encoder/n_layer");
    D.n_audio_head = (int)read_scalar(dir, "encoder/block_0/attn/n_head");
    read_tensor(dir, "encoder/positional_embedding", 2, sh, vals);
    D.n_audio_ctx = (int)sh[0];
    D.n_text_layer = (int)read_scalar(dir, "decoder/n_layer");
    D.n_text_head = (int)read_scalar(dir, "decoder/block_0/attn/n_head");
    read_tensor(dir, "decoder/token_embedding/weight", 2, sh, vals);
    D.n_vocab = (int)sh[0];
    D.n_text_state = (int)sh[1];
    read_tensor(dir, "decoder/positional_embedding", 2, sh, vals);
    D.n_text_ctx = (int)sh[0];
}

// every tensor file of the tree -> model_set_tensor (same order of traversal as load.rs:55-310)
void npy_tree_load(Model& m, const std::string& dir) {
    const wb_dims& D = m.dims;
    std::vector<int64_t> sh;
    std::vector<float> vals;
    auto tensor = [&](const std::string& path, int rank) {
        read_tensor(dir, path, rank, sh, vals);
        model_set_tensor(m, path.c_str(), vals.data(), sh.data(), rank);
    };
    auto scalar = [&](const std::string& path) {
        const float v = read_scalar(dir, path);
        const int64_t one = 1;
        model_set_tensor(m, path.c_str(), &v, &one, 1);
    };
    auto linear = [&](const std::string& p, bool bias) { tensor(p + "/weight", 2); if (bias) tensor(p + "/bias", 1); };
    auto ln = [&](const std::string& p) { tensor(p + "/weight", 1); tensor(p + "/bias", 1); scalar(p + "/eps"); };
    auto attn = [&](const std::string& p) { linear(p + "/query", true); linear(p + "/key", false); linear(p + "/value", true); linear(p + "/out", true); };
    auto block = [&](const std::string& p, bool cross) {
        attn(p + "/attn"); ln(p + "/attn_ln");
        if (cross) { attn(p + "/cross_attn"); ln(p + "/cross_attn_ln"); }
        linear(p + "/mlp/mlp1", true); linear(p + "/mlp/mlp2", true); ln(p + "/mlp_ln");
    };
    tensor("encoder/conv1/weight", 3); tensor("encoder/conv1/bias", 1);
    tensor("encoder/conv2/weight", 3); tensor("encoder/conv2/bias", 1);
    tensor("encoder/positional_embedding", 2);
    for (int i = 0; i < D.n_audio_layer; ++i) block("encoder/block_" + std::to_string(i), false);
    ln("encoder/ln_post");
    tensor("decoder/token_embedding/weight", 2);
    tensor("decoder/positional_embedding", 2);
    for (int i = 0; i < D.n_text_layer; ++i) block("decoder/block_" + std::to_string(i), true);
    ln("decoder/ln");
}

}  // namespace wb
