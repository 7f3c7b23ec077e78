// disk_replay.cpp — replay loader for the reference's on-disk rollout format.
//
// Reader for what DiskRollouts writes (reference src/rollouts_to_disk.jl:23-132) and DiskDataset reads
// (src/dataset.jl:1-82):
//   <root>/trajectory.csv     header `sample_names,selected_actions,selected_action_probabilities,returns`
//                             (after write_returns_to_disk, :106-132) or `...,rewards,terminal` (as written by
//                             update!, :34-40, 73-95)
//   <root>/states/sample_i.bson   BSON.@save of `state` (:47-51): BSON.jl lowers a bits-type array to
//                             {tag:"array", type:{tag:"datatype", params:[], name:[..]}, size:[..], data:<bytes>}
//                             (layout confirmed by the reference's own sample_1.bson) and a struct such as
//                             StateData(vertex_score, action_mask) (test/quad_game_utilities.jl:17-20) to
//                             {tag:"struct", type:{..}, data:[field...]}.
// The loader parses the CSV, reads the state files with a small thread pool (the format is one file per
// transition), converts Int64/Float64/Float32 arrays to the buffer's Float32 layout and appends to the device
// SoA buffer in chunks through ppo_buffer_append.  File I/O stays on the host (it is the reference's format);
// everything downstream is the device path.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ppo {
namespace {

// ---------------------------------------------------------------------------------------------
// minimal BSON reader
// ---------------------------------------------------------------------------------------------
struct BsonArray {          // a lowered Julia array of a bits type
    std::string eltype;     // "Int64", "Float32", "Float64", ...
    std::vector<int64_t> size;
    const uint8_t* data = nullptr;
    size_t bytes = 0;
    int64_t count() const { int64_t c = 1; for (int64_t s : size) c *= s; return c; }
};

struct Cursor {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    int32_t i32() { if (p + 4 > end) { ok = false; return 0; } int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    int64_t i64() { if (p + 8 > end) { ok = false; return 0; } int64_t v; memcpy(&v, p, 8); p += 8; return v; }
    std::string cstr() {
        const uint8_t* s = p;
        while (p < end && *p) ++p;
        if (p >= end) { ok = false; return ""; }
        std::string r((const char*)s, (size_t)(p - s)); ++p; return r;
    }
};

// skip one element value of BSON type `t`
bool skip_value(Cursor& c, uint8_t t) {
    switch (t) {
        case 0x01: c.p += 8; break;                               // double
        case 0x02: { int32_t n = c.i32(); c.p += n; break; }       // string
        case 0x03: case 0x04: { int32_t n; if (c.p + 4 > c.end) return false; memcpy(&n, c.p, 4); c.p += n; break; }
        case 0x05: { int32_t n = c.i32(); c.p += 1 + n; break; }   // binary
        case 0x08: c.p += 1; break;                               // bool
        case 0x0A: break;                                         // null
        case 0x10: c.p += 4; break;
        case 0x12: case 0x09: case 0x11: c.p += 8; break;
        default: return false;
    }
    return c.ok && c.p <= c.end;
}

// parse a document that is a lowered array; returns false when it is something else
bool parse_array_doc(const uint8_t* doc, const uint8_t* end, BsonArray& out);

// collect, in document order, every lowered bits-type array found under `doc` (recursing through struct data)
bool collect_arrays(const uint8_t* doc, const uint8_t* end, std::vector<BsonArray>& out, int depth = 0) {
    if (depth > 8 || doc + 5 > end) return false;
    Cursor c{doc, end};
    const int32_t len = c.i32();
    if (len < 5 || doc + len > end) return false;
    const uint8_t* dend = doc + len;
    // first pass: is this document itself a lowered array?
    BsonArray a;
    if (parse_array_doc(doc, dend, a)) { out.push_back(a); return true; }
    Cursor it{doc + 4, dend};
    while (it.p < dend && *it.p) {
        const uint8_t t = *it.p++;
        it.cstr();
        if (!it.ok) return false;
        if (t == 0x03 || t == 0x04) {
            collect_arrays(it.p, dend, out, depth + 1);
        }
        if (!skip_value(it, t)) return false;
    }
    return true;
}

bool parse_array_doc(const uint8_t* doc, const uint8_t* dend, BsonArray& out) {
    Cursor it{doc + 4, dend};
    bool is_array = false, have_data = false, have_size = false;
    while (it.p < dend && *it.p) {
        const uint8_t t = *it.p++;
        const std::string key = it.cstr();
        if (!it.ok) return false;
        if (key == "tag" && t == 0x02) {
            Cursor v{it.p, dend};
            const int32_t n = v.i32();
            if (n == 6 && memcmp(v.p, "array", 5) == 0) is_array = true;
        } else if (key == "type" && t == 0x03) {
            // {tag:"datatype", params:[], name:[module..., typename]} -> last element of `name`
            Cursor tdoc{it.p, dend};
            const int32_t tl = tdoc.i32();
            const uint8_t* tend = it.p + tl;
            while (tdoc.p < tend && *tdoc.p) {
                const uint8_t tt = *tdoc.p++;
                const std::string tk = tdoc.cstr();
                if (tk == "name" && tt == 0x04) {
                    Cursor nd{tdoc.p, tend};
                    const int32_t nl = nd.i32();
                    const uint8_t* nend = tdoc.p + nl;
                    while (nd.p < nend && *nd.p) {
                        const uint8_t nt = *nd.p++;
                        nd.cstr();
                        if (nt == 0x02) { Cursor sv{nd.p, nend}; const int32_t sl = sv.i32(); out.eltype.assign((const char*)sv.p, (size_t)std::max(0, sl - 1)); }
                        if (!skip_value(nd, nt)) return false;
                    }
                }
                if (!skip_value(tdoc, tt)) return false;
            }
        } else if (key == "size" && t == 0x04) {
            Cursor sd{it.p, dend};
            const int32_t sl = sd.i32();
            const uint8_t* send = it.p + sl;
            out.size.clear();
            while (sd.p < send && *sd.p) {
                const uint8_t st = *sd.p++;
                sd.cstr();
                if (st == 0x12) { Cursor v{sd.p, send}; out.size.push_back(v.i64()); }
                else if (st == 0x10) { Cursor v{sd.p, send}; out.size.push_back(v.i32()); }
                if (!skip_value(sd, st)) return false;
            }
            have_size = true;
        } else if (key == "data" && t == 0x05) {
            Cursor v{it.p, dend};
            const int32_t n = v.i32();
            out.data = v.p + 1;   // skip the subtype byte
            out.bytes = (size_t)n;
            have_data = true;
        }
        if (!skip_value(it, t)) return false;
    }
    return is_array && have_data && have_size;
}

size_t elsize(const std::string& t) {
    if (t == "Int64" || t == "Float64" || t == "UInt64") return 8;
    if (t == "Int32" || t == "Float32" || t == "UInt32") return 4;
    if (t == "Int16" || t == "UInt16" || t == "Float16") return 2;
    if (t == "Int8" || t == "UInt8" || t == "Bool") return 1;
    return 0;
}

// convert a lowered array to float32 (the byte order of a column-major Julia array is kept)
bool to_f32(const BsonArray& a, float* dst, int64_t expect) {
    const int64_t n = a.count();
    const size_t es = elsize(a.eltype);
    if (n != expect || es == 0 || (size_t)n * es != a.bytes) return false;
    if (a.eltype == "Float32") { memcpy(dst, a.data, (size_t)n * 4); return true; }
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* p = a.data + (size_t)i * es;
        if (a.eltype == "Int64") { int64_t v; memcpy(&v, p, 8); dst[i] = (float)v; }
        else if (a.eltype == "Float64") { double v; memcpy(&v, p, 8); dst[i] = (float)v; }
        else if (a.eltype == "Int32") { int32_t v; memcpy(&v, p, 4); dst[i] = (float)v; }
        else if (a.eltype == "UInt8" || a.eltype == "Bool") dst[i] = (float)p[0];
        else if (a.eltype == "Int8") dst[i] = (float)(int8_t)p[0];
        else return false;
    }
    return true;
}

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

std::vector<std::string> split_csv(const std::string& line) {
    std::vector<std::string> out;
    std::string cur; bool q = false;
    for (char ch : line) {
        if (ch == '"') q = !q;
        else if (ch == ',' && !q) { out.push_back(cur); cur.clear(); }
        else if (ch != '\r') cur.push_back(ch);
    }
    out.push_back(cur);
    return out;
}

}  // namespace
}  // namespace ppo

using namespace ppo;

extern "C" int ppo_bson_state_arrays(const char* path, int max_arrays, char* eltypes /* [max][16] */, int64_t* counts,
                                     int* ndims, int64_t* dims /* [max][4] */, int* n_arrays) {
    PPO_REQUIRE(path && eltypes && counts && ndims && dims && n_arrays, "bson_state_arrays: null argument");
    std::vector<uint8_t> raw;
    PPO_REQUIRE(read_file(path, raw), "cannot read %s", path);
    std::vector<BsonArray> arrays;
    PPO_REQUIRE(collect_arrays(raw.data(), raw.data() + raw.size(), arrays), "%s: not a BSON document", path);
    *n_arrays = (int)arrays.size();
    for (int i = 0; i < (int)arrays.size() && i < max_arrays; ++i) {
        snprintf(eltypes + 16 * i, 16, "%s", arrays[i].eltype.c_str());
        counts[i] = arrays[i].count();
        ndims[i] = (int)arrays[i].size.size();
        for (int d = 0; d < 4; ++d) dims[4 * i + d] = d < (int)arrays[i].size.size() ? arrays[i].size[d] : 0;
    }
    return PPO_OK;
}

extern "C" int ppo_disk_dataset_load(ppo_buf* buf, const char* root_directory, const char* trajectory_filename,
                                     const char* states_dirname, int n_threads, int64_t* n_loaded, int* has_returns) {
    PPO_REQUIRE(buf && root_directory, "disk_dataset_load: null argument");
    const std::string root(root_directory);
    const std::string csv_path = root + "/" + (trajectory_filename ? trajectory_filename : "trajectory.csv");
    const std::string states = root + "/" + (states_dirname ? states_dirname : "states");
    std::ifstream in(csv_path);
    PPO_REQUIRE(in.good(), "disk_dataset_load: cannot open %s", csv_path.c_str());   // @assert isfile(...), dataset.jl:7
    std::string line;
    PPO_REQUIRE((bool)std::getline(in, line), "disk_dataset_load: empty %s", csv_path.c_str());
    const std::vector<std::string> header = split_csv(line);
    int c_name = -1, c_act = -1, c_prob = -1, c_ret = -1, c_rew = -1, c_term = -1;
    for (int i = 0; i < (int)header.size(); ++i) {
        if (header[i] == "sample_names") c_name = i;
        else if (header[i] == "selected_actions") c_act = i;
        else if (header[i] == "selected_action_probabilities") c_prob = i;
        else if (header[i] == "returns") c_ret = i;
        else if (header[i] == "rewards") c_rew = i;
        else if (header[i] == "terminal") c_term = i;
    }
    PPO_REQUIRE(c_name >= 0 && c_act >= 0 && c_prob >= 0 && (c_ret >= 0 || (c_rew >= 0 && c_term >= 0)),
                "disk_dataset_load: %s has neither the returns schema nor the rewards/terminal schema", csv_path.c_str());
    const bool returns_schema = c_ret >= 0;
    std::vector<std::string> names;
    std::vector<int64_t> act;
    std::vector<float> prob, val;
    std::vector<uint8_t> term;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        const std::vector<std::string> f = split_csv(line);
        PPO_REQUIRE((int)f.size() >= (int)header.size(), "disk_dataset_load: malformed row %zu", names.size() + 2);
        names.push_back(f[c_name]);
        act.push_back(strtoll(f[c_act].c_str(), nullptr, 10));
        prob.push_back(strtof(f[c_prob].c_str(), nullptr));
        val.push_back(strtof(f[returns_schema ? c_ret : c_rew].c_str(), nullptr));
        term.push_back(returns_schema ? 0 : (uint8_t)(f[c_term] == "true" || f[c_term] == "1"));
    }
    const int64_t n = (int64_t)names.size();
    const int64_t fe = (int64_t)buf->nf * buf->nhe, A = buf->A;
    const int64_t chunk = 8192;
    std::vector<float> feat((size_t)chunk * fe), mask((size_t)chunk * A);
    if (n_threads < 1) n_threads = 1;
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t cn = std::min(chunk, n - c0);
        std::atomic<int64_t> next(0);
        std::atomic<int> failed(0);
        std::string first_error;
        auto work = [&]() {
            std::vector<uint8_t> raw;
            std::vector<BsonArray> arrays;
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= cn || failed.load()) return;
                const std::string path = states + "/" + names[(size_t)(c0 + i)];
                arrays.clear();
                bool ok = read_file(path, raw) && collect_arrays(raw.data(), raw.data() + raw.size(), arrays);
                // StateData: first array with fe elements -> vertex_score, first other array with A elements -> mask
                const BsonArray *vs = nullptr, *am = nullptr;
                if (ok) {
                    for (const BsonArray& a : arrays) {
                        if (!vs && a.count() == fe) vs = &a;
                        else if (!am && a.count() == A) am = &a;
                    }
                    ok = vs && am && to_f32(*vs, feat.data() + (size_t)i * fe, fe) && to_f32(*am, mask.data() + (size_t)i * A, A);
                }
                if (!ok && !failed.exchange(1)) first_error = path;
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        PPO_REQUIRE(!failed.load(), "disk_dataset_load: cannot read a StateData{[%d,%d],[%d]} state from %s", buf->nf, buf->nhe,
                    (int)A, first_error.c_str());
        PPO_TRY(ppo_buffer_append(buf, cn, feat.data(), mask.data(), act.data() + c0, prob.data() + c0, val.data() + c0,
                                  term.data() + c0));
    }
    if (n_loaded) *n_loaded = n;
    if (has_returns) *has_returns = returns_schema ? 1 : 0;
    return PPO_OK;
}
