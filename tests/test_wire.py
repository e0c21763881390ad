"""Wire codecs of the search call: api.SearchRequest / SearchResponse as msgpack and JSON
(src/api.zig:14-72, src/server.zig:84-142, 189-196) and the legacy line protocol's formats (src/legacy.zig:185-210,
286-296).  Encodings are cross-checked with python's json / msgpack; the reference's own e2e tests pin the key names
(tests/test_content_negotiation.py:20-34: r / i / s; tests/test_fingerprint_api.py: "query", "results", "id", "score";
tests/test_legacy.py:61-69: "1001:3 1002:2")."""
import json

import msgpack
import pytest

from _helpers import pkg


def test_json_request_defaults_and_clamps():
    r = pkg.decode_search_request(b'{"query": [101, 201, 301]}')
    assert (r.query, r.timeout, r.limit, r.min_score, r.score_pct) == ([101, 201, 301], 500, 40, None, 10)   # api.zig:14-22
    r = pkg.decode_search_request(json.dumps({"query": [4294967295, 0], "timeout": 99999, "limit": 5000, "min_score": 7,
                                              "score_pct": 55}).encode())
    assert (r.query, r.timeout, r.limit, r.min_score, r.score_pct) == ([4294967295, 0], 10000, 100, 7, 55)   # server.zig:192-193
    assert pkg.decode_search_request(b' { "limit" : 0 , "query" : [ ] , "min_score" : null } ').limit == 1
    for bad in (b'', b'[]', b'{"limit": 3}', b'{"query": [1], "extra": 1}', b'{"query": [4294967296]}', b'{"query": [-1]}',
                b'{"query": [1.5]}', b'{"query": [1]} x', b'{"query": [01]}', b'{"query": [1,]}', b'{"query": "1"}'):
        with pytest.raises(pkg.FpxError) as e:
            pkg.decode_search_request(bad)
        assert e.value.status == 2, bad                                              # error.BadRequest


def test_msgpack_request():
    r = pkg.decode_search_request(msgpack.packb({"q": [101, 201, 301]}), pkg.WIRE_MSGPACK)
    assert (r.query, r.timeout, r.limit, r.min_score, r.score_pct) == ([101, 201, 301], 500, 40, None, 10)
    r = pkg.decode_search_request(msgpack.packb({"q": [2 ** 32 - 1], "t": 0, "l": 7, "m": None, "s": 0}), pkg.WIRE_MSGPACK)
    assert (r.query, r.timeout, r.limit, r.min_score, r.score_pct) == ([2 ** 32 - 1], 0, 7, None, 0)
    r = pkg.decode_search_request(msgpack.packb({"m": 3, "q": list(range(300))}), pkg.WIRE_MSGPACK)
    assert r.min_score == 3 and r.query == list(range(300))
    for bad in (b'', msgpack.packb([1]), msgpack.packb({"t": 1}), msgpack.packb({"q": [1], "x": 1}), msgpack.packb({"q": [2 ** 32]}),
                msgpack.packb({"q": [-1]}), msgpack.packb({"q": [1]})[:-1], msgpack.packb({1: [1]})):
        with pytest.raises(pkg.FpxError):
            pkg.decode_search_request(bad, pkg.WIRE_MSGPACK)


def test_responses():
    res = [(1, 3), (2, 3), (4294967295, 1)]
    assert json.loads(pkg.encode_search_response(res)) == {"results": [{"id": i, "score": s} for i, s in res]}
    assert msgpack.unpackb(pkg.encode_search_response(res, pkg.WIRE_MSGPACK)) == {"r": [{"i": i, "s": s} for i, s in res]}
    assert pkg.encode_search_response(res, pkg.WIRE_MSGPACK) == msgpack.packb({"r": [{"i": i, "s": s} for i, s in res]})
    assert json.loads(pkg.encode_search_response([])) == {"results": []}
    assert msgpack.unpackb(pkg.encode_search_response([], pkg.WIRE_MSGPACK)) == {"r": []}
    many = [(i + 1, 100 - i % 7) for i in range(100)]
    assert msgpack.unpackb(pkg.encode_search_response(many, pkg.WIRE_MSGPACK)) == {"r": [{"i": i, "s": s} for i, s in many]}


def test_legacy_formats():
    assert pkg.legacy_parse_fingerprint("11000,12000,13000") == [11000, 12000, 13000]            # tests/test_legacy.py:61
    assert pkg.legacy_parse_fingerprint("-1,-2147483648,2147483647,+5") == [0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 5]
    assert pkg.legacy_parse_fingerprint("4294967296,9223372036854775807,-9223372036854775808") == [0, 0xFFFFFFFF, 0]  # @truncate
    for bad in ("", "1,,2", "1,", "a", "1 ,2", "9223372036854775808", "-9223372036854775809", "-"):
        with pytest.raises(pkg.FpxError):
            pkg.legacy_parse_fingerprint(bad)
    assert pkg.legacy_format_results([(1001, 3), (1002, 2)]) == "1001:3 1002:2"                   # tests/test_legacy.py:69
    assert pkg.legacy_format_results([]) == ""
