// fpx_wire.cu — wire codecs of the search call, for a search-only endpoint without the Zig host.
//
//   request   api.SearchRequest (src/api.zig:14-27): msgpack map with one-letter keys q, t, l, m, s; JSON object
//             with the field names query, timeout, limit, min_score, score_pct (std.json, unknown fields rejected);
//             defaults timeout 500, limit 40, min_score null, score_pct 10; src/server.zig:189-194 clamps limit to
//             [1, 100] and timeout to <= 10000
//   response  api.SearchResponse (src/api.zig:56-72): {"r": [{"i": id, "s": score}]} / {"results": [{"id":, "score":}]}
//   legacy    "search <csv of signed decimals>" -> u32 hashes (src/legacy.zig:286-296), reply "id:score id:score ..."
//             (src/legacy.zig:203-208)
// The msgpack reader accepts any valid encoding (the reference's encoder is not vendored).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fpx.h"
#include "fpx_filefmt.h"

namespace fpx {
void set_last_error(const std::string &msg); // fpx_api.cu
}

namespace {

fpx_status bad(const char *m) {
    fpx::set_last_error(m);
    return FPX_INVALID_ARGUMENT; // error.BadRequest -> 400 (server.zig:111-113)
}

struct Json {
    const char *p, *end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    bool lit(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    bool key(std::string &k) { // a plain string without escapes (the field names have none)
        ws();
        if (p >= end || *p != '"') return false;
        const char *q = ++p;
        while (p < end && *p != '"' && *p != '\\') ++p;
        if (p >= end || *p != '"') return false;
        k.assign(q, p);
        ++p;
        return true;
    }
    bool uint(uint64_t &v, uint64_t max) { // a non-negative integer literal
        ws();
        if (p >= end || *p < '0' || *p > '9') return false;
        if (*p == '0' && p + 1 < end && p[1] >= '0' && p[1] <= '9') return false; // no leading zeros in JSON
        v = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            v = v * 10 + (uint64_t)(*p - '0');
            if (v > max) return false;
            ++p;
        }
        if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) return false;
        return true;
    }
    bool null() {
        ws();
        if (end - p >= 4 && std::memcmp(p, "null", 4) == 0) {
            p += 4;
            return true;
        }
        return false;
    }
};

} // namespace

extern "C" {

fpx_status fpx_wire_decode_search_request(uint32_t format, const uint8_t *data, uint64_t size, fpx_wire_search_request *out) {
    if (!out || (size && !data)) return bad("null argument");
    std::vector<uint32_t> q;
    out->query = nullptr;
    out->n_terms = 0;
    out->timeout = 500;  // api.zig:7
    out->limit = 40;     // api.zig:9
    out->has_min_score = 0;
    out->min_score = 0;
    out->score_pct = 10; // api.zig:22
    bool have_query = false;
    try {
        if (format == FPX_WIRE_MSGPACK) {
            fpx::MsgpackReader r(data, (size_t)size);
            uint64_t n = 0;
            if (!r.read_map(n)) return bad("request is not a msgpack map");
            for (uint64_t i = 0; i < n; ++i) {
                std::string k;
                uint64_t v = 0;
                if (!r.read_str(k)) return bad("request key is not a string");
                if (k == "q") {
                    uint64_t m = 0;
                    if (!r.read_array(m) || m > (uint64_t)(r.end - r.p)) return bad("query is not an array");
                    q.reserve((size_t)m);
                    for (uint64_t j = 0; j < m; ++j) {
                        if (!r.read_uint(v) || v > 0xFFFFFFFFull) return bad("query term is not a u32");
                        q.push_back((uint32_t)v);
                    }
                    have_query = true;
                } else if (k == "t") {
                    if (!r.read_uint(v) || v > 0xFFFFFFFFull) return bad("timeout is not a u32");
                    out->timeout = (uint32_t)v;
                } else if (k == "l") {
                    if (!r.read_uint(v) || v > 0xFFFFFFFFull) return bad("limit is not a u32");
                    out->limit = (uint32_t)v;
                } else if (k == "m") {
                    if (!r.read_nil()) {
                        if (!r.read_uint(v) || v > 0xFFFFFFFFull) return bad("min_score is not a u32");
                        out->has_min_score = 1;
                        out->min_score = (uint32_t)v;
                    }
                } else if (k == "s") {
                    if (!r.read_uint(v) || v > 0xFFFFFFFFull) return bad("score_pct is not a u32");
                    out->score_pct = (uint32_t)v;
                } else {
                    return bad("unknown request field");
                }
            }
            if (!r.ok) return bad("truncated msgpack request");
        } else if (format == FPX_WIRE_JSON) {
            Json j{reinterpret_cast<const char *>(data), reinterpret_cast<const char *>(data) + size};
            if (!j.lit('{')) return bad("request is not a JSON object");
            bool first = true;
            while (!j.lit('}')) {
                if (!first && !j.lit(',')) return bad("malformed JSON object");
                first = false;
                std::string k;
                uint64_t v = 0;
                if (!j.key(k) || !j.lit(':')) return bad("malformed JSON object");
                if (k == "query") {
                    if (!j.lit('[')) return bad("query is not an array");
                    bool f2 = true;
                    while (!j.lit(']')) {
                        if (!f2 && !j.lit(',')) return bad("malformed query array");
                        f2 = false;
                        if (!j.uint(v, 0xFFFFFFFFull)) return bad("query term is not a u32");
                        q.push_back((uint32_t)v);
                    }
                    have_query = true;
                } else if (k == "timeout") {
                    if (!j.uint(v, 0xFFFFFFFFull)) return bad("timeout is not a u32");
                    out->timeout = (uint32_t)v;
                } else if (k == "limit") {
                    if (!j.uint(v, 0xFFFFFFFFull)) return bad("limit is not a u32");
                    out->limit = (uint32_t)v;
                } else if (k == "min_score") {
                    if (!j.null()) {
                        if (!j.uint(v, 0xFFFFFFFFull)) return bad("min_score is not a u32");
                        out->has_min_score = 1;
                        out->min_score = (uint32_t)v;
                    }
                } else if (k == "score_pct") {
                    if (!j.uint(v, 0xFFFFFFFFull)) return bad("score_pct is not a u32");
                    out->score_pct = (uint32_t)v;
                } else {
                    return bad("unknown request field"); // std.json: ignore_unknown_fields = false
                }
            }
            j.ws();
            if (j.p != j.end) return bad("trailing characters after the JSON object");
        } else {
            return bad("unknown wire format");
        }
        if (!have_query) return bad("missing field: query"); // the only field without a default
        // server.zig:192-193: sanitize untrusted values
        out->limit = std::max(std::min(out->limit, 100u), 1u);
        out->timeout = std::min(out->timeout, 10000u);
        uint32_t *mem = static_cast<uint32_t *>(std::malloc(std::max<size_t>(q.size(), 1) * sizeof(uint32_t)));
        if (!mem) return FPX_OUT_OF_MEMORY;
        if (!q.empty()) std::memcpy(mem, q.data(), q.size() * sizeof(uint32_t));
        out->query = mem;
        out->n_terms = q.size();
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    return FPX_OK;
}

fpx_status fpx_wire_encode_search_response(uint32_t format, const uint32_t *ids, const uint32_t *scores, uint32_t n,
                                           uint8_t **out, uint64_t *out_size) {
    if (!out || !out_size || (n && (!ids || !scores))) return bad("null argument");
    try {
        std::vector<uint8_t> b;
        if (format == FPX_WIRE_MSGPACK) {
            fpx::MsgpackWriter w(b);
            w.map(1);
            w.str("r");
            w.array(n);
            for (uint32_t i = 0; i < n; ++i) {
                w.map(2);
                w.str("i"), w.uint(ids[i]);
                w.str("s"), w.uint(scores[i]);
            }
        } else if (format == FPX_WIRE_JSON) {
            std::string s = "{\"results\":[";
            char tmp[64];
            for (uint32_t i = 0; i < n; ++i) {
                std::snprintf(tmp, sizeof tmp, "%s{\"id\":%u,\"score\":%u}", i ? "," : "", ids[i], scores[i]);
                s += tmp;
            }
            s += "]}";
            b.assign(s.begin(), s.end());
        } else {
            return bad("unknown wire format");
        }
        uint8_t *mem = static_cast<uint8_t *>(std::malloc(std::max<size_t>(b.size(), 1)));
        if (!mem) return FPX_OUT_OF_MEMORY;
        std::memcpy(mem, b.data(), b.size());
        *out = mem;
        *out_size = b.size();
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    return FPX_OK;
}

// legacy.zig:286-296: comma-separated signed decimals (i64), each truncated to u32
fpx_status fpx_legacy_parse_fingerprint(const char *text, uint64_t len, uint32_t **out_terms, uint64_t *out_n) {
    if (!out_terms || !out_n || (len && !text)) return bad("null argument");
    *out_terms = nullptr;
    *out_n = 0;
    if (len == 0) return bad("empty fingerprint");
    std::vector<uint32_t> q;
    try {
        uint64_t pos = 0;
        while (pos <= len) {
            uint64_t end = pos;
            while (end < len && text[end] != ',') ++end;
            // std.fmt.parseInt(i64, tok, 10): optional sign, digits (underscores are not produced by clients)
            uint64_t i = pos;
            bool neg = false;
            if (i < end && (text[i] == '-' || text[i] == '+')) neg = text[i++] == '-';
            if (i == end) return bad("invalid fingerprint");
            unsigned long long mag = 0;
            for (; i < end; ++i) {
                if (text[i] < '0' || text[i] > '9') return bad("invalid fingerprint");
                if (mag > (0x8000000000000000ull - (unsigned)(text[i] - '0')) / 10) return bad("invalid fingerprint");
                mag = mag * 10 + (unsigned)(text[i] - '0');
            }
            if (!neg && mag > 0x7FFFFFFFFFFFFFFFull) return bad("invalid fingerprint");
            const unsigned long long bits = neg ? (0ull - mag) : mag; // two's complement of the i64
            q.push_back((uint32_t)bits);
            pos = end + 1;
            if (end == len) break;
        }
        uint32_t *mem = static_cast<uint32_t *>(std::malloc(std::max<size_t>(q.size(), 1) * sizeof(uint32_t)));
        if (!mem) return FPX_OUT_OF_MEMORY;
        std::memcpy(mem, q.data(), q.size() * sizeof(uint32_t));
        *out_terms = mem;
        *out_n = q.size();
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    return FPX_OK;
}

// legacy.zig:203-208: "id:score id:score ..."
fpx_status fpx_legacy_format_results(const uint32_t *ids, const uint32_t *scores, uint32_t n, uint8_t **out, uint64_t *out_size) {
    if (!out || !out_size || (n && (!ids || !scores))) return bad("null argument");
    try { // no exception may cross the C boundary
        std::string s;
        char tmp[32];
        for (uint32_t i = 0; i < n; ++i) {
            std::snprintf(tmp, sizeof tmp, "%s%u:%u", i ? " " : "", ids[i], scores[i]);
            s += tmp;
        }
        uint8_t *mem = static_cast<uint8_t *>(std::malloc(std::max<size_t>(s.size(), 1)));
        if (!mem) return FPX_OUT_OF_MEMORY;
        std::memcpy(mem, s.data(), s.size());
        *out = mem;
        *out_size = s.size();
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    return FPX_OK;
}

void fpx_wire_free(void *p) { std::free(p); }

} // extern "C"
