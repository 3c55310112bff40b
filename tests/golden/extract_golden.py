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


if __name__ == "__main__":
    main()
