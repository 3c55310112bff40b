#!/usr/bin/env python3
"""Extract the golden vectors the reference's own tests hold for the hot path.

Run in the authoring container (needs /root/reference; the GPU box does not have it):

    python tests/golden/extract_golden.py

It parses (never imports/executes) the reference's Python kernel tests with ``ast`` and writes the
literal arguments of their known-answer helper calls to JSON fixtures next to this file:

    conv_ops.json     tensorflow/python/kernel_tests/conv_ops_test.py
                      Conv2DTest._VerifyValues / _RunAndVerifyBackpropInput / ...Filter
    pooling_ops.json  tensorflow/python/kernel_tests/pooling_ops_test.py
                      PoolingTest._VerifyValues(nn_ops.max_pool, ...) and _testMaxPoolGradDirect

    half_plus_two.graph_def.pb / half_plus_two.nodes.json
                      tensorflow/cc/saved_model/testdata/half_plus_two*/00000123: the GraphDef a real
                      TensorFlow 1.0 serialized (bytes cut out of saved_model.pb: SavedModel.meta_graphs[0]
                      .graph_def) and the same nodes read from the text-format twin saved_model.pbtxt --
                      the golden pair for the hand-written wire reader (graph_def_wire.cc)

Every record carries the source file and line of the call it came from.  Inputs of these tests
are "incrementing numbers from 1" in row-major order (conv_ops_test.py:216-219,
pooling_ops_test.py:128-131) unless explicit input lists are given.
"""
import ast
import json
import os
import sys

REF = "/root/reference/tensorflow/python/kernel_tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def _literal(node, env):
    """literal_eval with lookup of names previously bound to literals in the same function."""
    if isinstance(node, ast.Name) and node.id in env:
        return env[node.id]
    return ast.literal_eval(node)


def _calls_in_function(fn):
    """Yield (call, env) for self.<helper>(...) calls; env = literal assignments seen so far."""
    env = {}
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(
                node.targets[0], ast.Name):
            try:
                env[node.targets[0].id] = ast.literal_eval(node.value)
            except (ValueError, SyntaxError):
                pass
    for node in ast.walk(fn):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(
                node.func.value, ast.Name) and node.func.value.id == "self":
            yield node, env


def _kwargs(call, env, wanted):
    out = {}
    for kw in call.keywords:
        if kw.arg in wanted:
            out[kw.arg] = _literal(kw.value, env)
    return out


def extract_conv():
    path = os.path.join(REF, "conv_ops_test.py")
    tree = ast.parse(open(path).read())
    records = []
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Conv2DTest"]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            for call, env in _calls_in_function(fn):
                helper = call.func.attr
                try:
                    if helper == "_VerifyValues":
                        kw = _kwargs(call, env, {"tensor_in_sizes", "filter_in_sizes", "strides",
                                                 "padding", "expected"})
                        if len(kw) != 5:
                            continue
                        kind = "conv2d"
                    elif helper in ("_RunAndVerifyBackpropInput", "_RunAndVerifyBackpropFilter"):
                        kw = _kwargs(call, env, {"input_sizes", "filter_sizes", "output_sizes",
                                                 "strides", "padding", "expected"})
                        if len(kw) != 6:
                            continue
                        kind = ("conv2d_backprop_input" if helper.endswith("Input")
                                else "conv2d_backprop_filter")
                    else:
                        continue
                except (ValueError, SyntaxError):
                    continue
                kw.update(kind=kind, test=fn.name,
                          source="tensorflow/python/kernel_tests/conv_ops_test.py:%d" % call.lineno)
                records.append(kw)
    return records


def extract_pooling():
    path = os.path.join(REF, "pooling_ops_test.py")
    tree = ast.parse(open(path).read())
    records = []
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PoolingTest"]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            for call, env in _calls_in_function(fn):
                helper = call.func.attr
                try:
                    if helper == "_VerifyValues":
                        # first positional arg selects the op: keep nn_ops.max_pool only
                        if not call.args or not isinstance(call.args[0], ast.Attribute) or \
                                call.args[0].attr != "max_pool":
                            continue
                        kw = _kwargs(call, env, {"input_sizes", "ksize", "strides", "padding",
                                                 "expected"})
                        if len(kw) != 5:
                            continue
                        kind = "max_pool"
                    elif helper == "_testMaxPoolGradDirect":
                        kw = _kwargs(call, env, {"input_sizes", "output_sizes", "window_rows",
                                                 "window_cols", "row_stride", "col_stride",
                                                 "padding"})
                        names = ["input_data", "output_backprop", "expected_input_backprop"]
                        for name, arg in zip(names, call.args):
                            kw[name] = _literal(arg, env)
                        if len(kw) != 10:
                            continue
                        kind = "max_pool_grad"
                    else:
                        continue
                except (ValueError, SyntaxError):
                    continue
                kw.update(kind=kind, test=fn.name,
                          source="tensorflow/python/kernel_tests/pooling_ops_test.py:%d"
                          % call.lineno)
                records.append(kw)
    return records


# ---------------------------------------------------------------- half_plus_two GraphDef golden
SAVED_MODEL = "/root/reference/tensorflow/cc/saved_model/testdata"
TYPES_PROTO = "/root/reference/tensorflow/core/framework/types.proto"


def _varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _field_bytes(buf, want):
    """The payload of the first length-delimited field `want` of a serialized message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            _, pos = _varint(buf, pos)
        elif wt == 1:
            pos += 8
        elif wt == 5:
            pos += 4
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if field == want:
                return buf[pos:pos + n]
            pos += n
        else:
            raise ValueError("unexpected wire type %d" % wt)
    raise KeyError(want)


def _parse_text_proto(text):
    """Protobuf text format -> list of (key, value); value is a str/number/bool or a nested list."""
    import re
    tokens = re.findall(r'"(?:[^"\\]|\\.)*"|[{}]|[^\s{}:]+:?', text)
    pos = 0

    def block():
        nonlocal pos
        items = []
        while pos < len(tokens) and tokens[pos] != "}":
            key = tokens[pos]
            pos += 1
            if key.endswith(":"):
                val = tokens[pos]
                pos += 1
                if val == "{":        # "key: {" is legal too
                    items.append((key[:-1], block()))
                    pos += 1
                else:
                    items.append((key[:-1], val))
            else:
                assert tokens[pos] == "{", tokens[pos - 1:pos + 2]
                pos += 1
                items.append((key, block()))
                pos += 1
        return items
    return block()


def _unquote(tok):
    body = tok[1:-1]
    out, i = bytearray(), 0
    while i < len(body):
        c = body[i]
        if c == "\\":
            n = body[i + 1]
            if n in "01234567":
                out.append(int(body[i + 1:i + 4], 8))
                i += 4
                continue
            out.append({"n": 10, "t": 9, "r": 13, '"': 34, "'": 39, "\\": 92}[n])
            i += 2
            continue
        out.extend(c.encode("utf-8"))
        i += 1
    return out.decode("latin-1")


def extract_half_plus_two():
    import re
    dtypes = dict((m.group(1), int(m.group(2))) for m in
                  re.finditer(r"^\s*(DT_[A-Z0-9_]+)\s*=\s*(\d+);", open(TYPES_PROTO).read(), re.M))
    pb = open(os.path.join(SAVED_MODEL, "half_plus_two/00000123/saved_model.pb"), "rb").read()
    graph_def = _field_bytes(_field_bytes(pb, 2), 2)  # SavedModel.meta_graphs[0].graph_def
    with open(os.path.join(HERE, "half_plus_two.graph_def.pb"), "wb") as f:
        f.write(graph_def)
    text = open(os.path.join(SAVED_MODEL, "half_plus_two_pbtxt/00000123/saved_model.pbtxt")).read()
    tree = _parse_text_proto(text)
    meta = dict(tree)["meta_graphs"]
    gd = [v for k, v in meta if k == "graph_def"][0]
    nodes = []
    for key, node in gd:
        if key != "node":
            continue
        rec = {"name": None, "op": None, "device": "", "input": [], "attr": {}}
        for k, v in node:
            if k in ("name", "op", "device"):
                rec[k] = _unquote(v)
            elif k == "input":
                rec["input"].append(_unquote(v))
            elif k == "attr":
                d = dict(v)
                value = d["value"]
                (vk, vv), = value if len(value) == 1 else [("none", None)]
                if vk == "type":
                    rec["attr"][_unquote(d["key"])] = {"type": dtypes[vv]}
                elif vk == "i":
                    rec["attr"][_unquote(d["key"])] = {"i": int(vv)}
                elif vk == "b":
                    rec["attr"][_unquote(d["key"])] = {"b": vv == "true"}
                elif vk == "s":
                    rec["attr"][_unquote(d["key"])] = {"s": _unquote(vv)}
                elif vk == "shape":
                    dims = [int(dict(dv).get("size", 0)) for dk, dv in vv if dk == "dim"]
                    unknown = any(dk == "unknown_rank" for dk, _ in vv) or any(x < 0 for x in dims)
                    rec["attr"][_unquote(d["key"])] = {"shape": None if unknown else dims}
                elif vk == "tensor":
                    td = dict(vv)
                    dims = [int(dict(dv).get("size", 0)) for dk, dv in td.get("tensor_shape", []) if dk == "dim"]
                    t = {"dtype": dtypes[td["dtype"]], "shape": dims}
                    if "float_val" in td:
                        t["float_val"] = [float(x) for k2, x in vv if k2 == "float_val"]
                    if "int_val" in td:
                        t["int_val"] = [int(x) for k2, x in vv if k2 == "int_val"]
                    rec["attr"][_unquote(d["key"])] = {"tensor": t}
                else:
                    rec["attr"][_unquote(d["key"])] = {"other": vk}
        nodes.append(rec)
    with open(os.path.join(HERE, "half_plus_two.nodes.json"), "w") as f:
        json.dump({"source": "tensorflow/cc/saved_model/testdata/half_plus_two_pbtxt/00000123/"
                             "saved_model.pbtxt (meta_graphs[0].graph_def)",
                   "known_answer": {"source": "tensorflow/cc/saved_model/loader_test.cc: y = 0.5 * x + 2",
                                    "x": [0.0, 1.0, 2.0, 3.0], "y": [2.0, 2.5, 3.0, 3.5]},
                   "nodes": nodes}, f, indent=1, sort_keys=True)
    print("half_plus_two: %d bytes of GraphDef, %d nodes" % (len(graph_def), len(nodes)))


def _sanitize(obj):
    """JSON has no NaN literal in strict mode: encode float('nan') as the string "nan"."""
    if isinstance(obj, float) and obj != obj:
        return "nan"
    if isinstance(obj, (list, tuple)):
        return [_sanitize(x) for x in obj]
    if isinstance(obj, dict):
        return {k: _sanitize(v) for k, v in obj.items()}
    return obj


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s (run this in the authoring container)" % REF)
    for name, recs in (("conv_ops.json", extract_conv()), ("pooling_ops.json", extract_pooling())):
        # de-duplicate (helpers are called once per data_format / use_gpu in loops)
        seen, uniq = set(), []
        for r in recs:
            key = json.dumps(_sanitize({k: v for k, v in r.items() if k != "source"}),
                             sort_keys=True)
            if key not in seen:
                seen.add(key)
                uniq.append(_sanitize(r))
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(uniq, f, indent=1, sort_keys=True)
        print("%s: %d records" % (name, len(uniq)))
    extract_half_plus_two()


if __name__ == "__main__":
    main()
