"""GraphDef wire-format reader / writer: a GraphDef serialized by a real
TensorFlow 1.0 (the reference's half_plus_two SavedModel test data, tests/golden/extract_golden.py)
is read by the hand-written codec and compared node by node with the text-format twin the reference
ships; graphs built here round-trip through the wire format; and (GPU) the imported real graph
runs to the reference's known answer y = 0.5 x + 2 (cc/saved_model/loader_test.cc:90)."""
import json
import os
import re

import numpy as np
import pytest

from simple_tensorflow_b200 import client
from simple_tensorflow_b200 import ops as tf

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_PB = os.path.join(HERE, "golden", "half_plus_two.graph_def.pb")
GOLDEN_NODES = os.path.join(HERE, "golden", "half_plus_two.nodes.json")


def _parse_dump(text):
    nodes = []
    for line in text.splitlines():
        kind, name, op, device, inputs, attrs = line.split("\t")
        assert kind == "node"
        amap = {}
        # attr summaries are separated by ';' -- but quoted strings may contain ';' too
        for m in re.finditer(r'([^=;]+)=((?:"(?:[^"\\]|\\.)*"|[^;])*)', attrs):
            amap[m.group(1)] = m.group(2)
        nodes.append({"name": name, "op": op, "device": device,
                      "input": inputs.split(",") if inputs else [], "attr": amap})
    return nodes


def test_reads_a_graphdef_written_by_real_tensorflow():
    golden = json.load(open(GOLDEN_NODES))["nodes"]
    got = _parse_dump(client.graph_def_to_text(open(GOLDEN_PB, "rb").read()))
    assert [n["name"] for n in got] == [n["name"] for n in golden]
    for g, w in zip(got, golden):
        assert (g["op"], g["device"], g["input"]) == (w["op"], w["device"], w["input"]), w["name"]
        assert set(g["attr"]) == set(w["attr"]), w["name"]
        for key, want in w["attr"].items():
            have = g["attr"][key]
            if "type" in want:
                assert have == "type:%d" % want["type"], (w["name"], key)
            elif "i" in want:
                assert have == "i:%d" % want["i"], (w["name"], key)
            elif "b" in want:
                assert have == ("b:true" if want["b"] else "b:false"), (w["name"], key)
            elif "s" in want and all(32 <= ord(c) < 127 and c not in '"\\' for c in want["s"]):
                assert have == 's:"%s"' % want["s"], (w["name"], key)
            elif "shape" in want:
                if want["shape"] is None:
                    assert have.startswith("raw:"), (w["name"], key)  # unknown dims: kept verbatim
                else:
                    assert have == "shape:[%s]" % ",".join(map(str, want["shape"])), (w["name"], key)
            elif "tensor" in want and want["tensor"]["dtype"] == 1 and "float_val" in want["tensor"]:
                t = want["tensor"]
                n = int(np.prod(t["shape"])) if t["shape"] else 1
                vals = (t["float_val"] + [t["float_val"][-1]] * n)[:min(n, 8)]
                assert have == "tensor:1:[%s]:%s" % (",".join(map(str, t["shape"])),
                                                      ",".join("%.9g" % v for v in vals)), (w["name"], key)
            elif "other" in want and want["other"] == "list":
                assert have.split(":")[0] in ("raw", "list_i", "list_s", "list_type"), (w["name"], key)


def test_real_graphdef_survives_a_parse_serialize_cycle():
    # nothing is lost: re-serialized bytes parse to the same dump (list(shape) attrs, string
    # tensors and the versions field ride along verbatim)
    original = open(GOLDEN_PB, "rb").read()
    tf.reset_default_graph()
    g = tf.get_default_graph()
    ops_by_name = g.import_graph_def(original)
    assert "y" in ops_by_name and ops_by_name["y"].type == "Add"
    assert ops_by_name["ParseExample/ParseExample"].type == "ParseExample"   # opaque, but imported
    again = g.as_graph_def()
    a, b = _parse_dump(client.graph_def_to_text(original)), _parse_dump(client.graph_def_to_text(again))
    assert [n["name"] for n in a] == [n["name"] for n in b]
    for x, y in zip(a, b):
        assert (x["op"], x["input"], x["device"]) == (y["op"], y["input"], y["device"])
        for k, v in x["attr"].items():   # import adds defaulted attrs, never drops or changes one
            assert y["attr"][k] == v, (x["name"], k)
    # and the export is a fixed point
    tf.reset_default_graph()
    g2 = tf.get_default_graph()
    g2.import_graph_def(again)
    assert g2.as_graph_def() == again


def test_graph_built_here_round_trips_and_is_byte_stable(rng):
    tf.reset_default_graph()
    x = tf.placeholder(tf.float32, [4, 3], "x")
    w = tf.Variable(rng.randn(3, 5).astype(np.float32), name="w")
    b = tf.constant(np.arange(5, dtype=np.float32), name="b")
    y = tf.relu(tf.bias_add(tf.matmul(x, w, transpose_b=False, name="mm"), b), name="y")
    idx = tf.argmax(y, 1)
    first = tf.get_default_graph().as_graph_def()
    text = _parse_dump(client.graph_def_to_text(first))
    by_name = {n["name"]: n for n in text}
    assert by_name["mm"]["op"] == "MatMul" and by_name["mm"]["attr"]["transpose_a"] == "b:false"
    assert by_name["b"]["attr"]["value"] == "tensor:1:[5]:0,1,2,3,4"
    assert by_name["x"]["attr"]["shape"] == "shape:[4,3]"
    assert by_name[idx.op.name]["input"][0] == "y"
    tf.reset_default_graph()
    g2 = tf.get_default_graph()
    imported = g2.import_graph_def(first, name="copy")
    assert imported["copy/mm"].type == "MatMul"
    # prefixing rewrites names and inputs (control inputs included) and nothing else
    second = _parse_dump(client.graph_def_to_text(g2.as_graph_def()))
    assert [n["name"] for n in second] == ["copy/" + n["name"] for n in text]
    for a, c in zip(text, second):
        assert c["input"] == [("^copy/" + i[1:]) if i.startswith("^") else "copy/" + i for i in a["input"]]
        assert a["attr"] == c["attr"]


def test_imported_operations_can_be_walked():
    # TF_OperationInput / GetControlInputs / GetAttr*: the importer wires inputs, control inputs,
    # attributes and the static shapes the nodes state themselves
    tf.reset_default_graph()
    g = tf.get_default_graph()
    ops = g.import_graph_def(open(GOLDEN_PB, "rb").read())
    y = ops["y"]
    assert [i.op.name for i in y.inputs] == ["Mul", "b/read"] and y.inputs[0].index == 0
    assert ops["x2"].inputs[0].op.name == "ParseExample/ParseExample" and ops["x2"].inputs[0].index == 1
    assert sorted(c.name for c in ops["init"].control_inputs) == ["a/Assign", "b/Assign", "c/Assign"]
    assert y.attrs["T"] == ("type", tf.float32)
    assign = ops["a/Assign"]
    assert assign.attrs["use_locking"] is True and assign.attrs["validate_shape"] is True
    assert ops["a"].attrs["shape"] == ("shape", []) and ops["a"].attrs["container"] == ""
    assert g.shapes["a:0"] == () and g.shapes["a/initial_value:0"] == ()
    assert ops["save/SaveV2"].type == "SaveV2"          # opaque node: still listed, still wired
    assert [i.op.name for i in ops["save/SaveV2"].inputs][-3:] == ["a", "b", "c"]
    # a graph built here and re-imported: attributes survive in the form create_op takes
    tf.reset_default_graph()
    x = tf.placeholder(tf.float32, [2, 8, 8, 4], "x")
    w = tf.Variable(np.zeros((3, 3, 4, 8), np.float32), name="w")
    tf.max_pool(tf.conv2d(x, w, [1, 2, 2, 1], "SAME", name="conv"), [1, 2, 2, 1], [1, 2, 2, 1], "VALID",
                name="pool")
    blob = tf.get_default_graph().as_graph_def()
    tf.reset_default_graph()
    again = tf.get_default_graph().import_graph_def(blob)
    conv = again["conv"]
    assert conv.attrs["strides"] == ("ints", [1, 2, 2, 1]) and conv.attrs["padding"] == "SAME"
    assert conv.attrs["data_format"] == "NHWC" and conv.attrs["use_cudnn_on_gpu"] is True
    assert again["pool"].attrs["ksize"] == ("ints", [1, 2, 2, 1])
    assert tf.get_default_graph().shapes["x:0"] == (2, 8, 8, 4)
    assert tf.get_default_graph().shapes["w:0"] == (3, 3, 4, 8)


def test_training_graph_on_top_of_an_imported_model():
    # forward model built, serialized, imported; gradients + SGD are then attached to the IMPORTED
    # nodes: same backward ops as on the original graph (shapes are re-inferred on import)
    def forward():
        tf.reset_default_graph()
        x = tf.placeholder(tf.float32, [8, 12, 12, 4], "x")
        lab = tf.placeholder(tf.float32, [8, 10], "lab")
        w = tf.Variable(np.zeros((3, 3, 4, 8), np.float32), name="w")
        b = tf.Variable(np.zeros(8, np.float32), name="b")
        wf = tf.Variable(np.zeros((6 * 6 * 8, 10), np.float32), name="wf")
        bf = tf.Variable(np.zeros(10, np.float32), name="bf")
        c = tf.relu(tf.bias_add(tf.conv2d(x, w, [1, 1, 1, 1], "SAME", name="conv"), b))
        p = tf.max_pool(c, [1, 2, 2, 1], [1, 2, 2, 1], "SAME", name="pool")
        logits = tf.bias_add(tf.matmul(tf.reshape(p, [8, 288], name="flat"), wf, name="fc"), bf)
        return tf.reduce_mean(tf.softmax_cross_entropy_with_logits(logits, lab), name="loss"), [w, b, wf, bf]

    loss, variables = forward()
    original_shapes = dict(tf.get_default_graph().shapes)
    blob = tf.get_default_graph().as_graph_def()
    tf.GradientDescentOptimizer(0.1).minimize(loss, variables)
    want = sorted(op.type for op in tf.get_default_graph().operations)

    tf.reset_default_graph()
    g = tf.get_default_graph()
    g.import_graph_def(blob)
    assert {k: g.shapes.get(k) for k in original_shapes} == original_shapes   # inferred again
    imported = [tf.Variable.from_imported(g.get_tensor_by_name(n + ":0")) for n in ("w", "b", "wf", "bf")]
    train = tf.GradientDescentOptimizer(0.1).minimize(g.get_tensor_by_name("loss:0"), imported)
    assert sorted(op.type for op in g.operations) == want and train.type == "NoOp"
    # new names never collide with imported ones ("Const_1" exists in the import)
    names = [op.name for op in g.operations]
    assert len(names) == len(set(names))
    with pytest.raises(ValueError):
        tf.Variable.from_imported(g.get_tensor_by_name("x:0"))


def test_import_errors():
    tf.reset_default_graph()
    g = tf.get_default_graph()
    with pytest.raises(client.OpError) as e:   # truncated protobuf
        g.import_graph_def(open(GOLDEN_PB, "rb").read()[:-3])
    assert e.value.error_code == 3
    tf.reset_default_graph()
    g = tf.get_default_graph()
    tf.placeholder(tf.float32, [1], "x")
    good = g.as_graph_def()
    with pytest.raises(client.OpError) as e:   # duplicate names without a prefix
        g.import_graph_def(good)
    assert e.value.error_code == 3 and "Duplicate" in e.value.message
    # an input that names no node: hand-made NodeDef{name:"r" op:"Relu" input:"nope" attr T=float}
    attr = b"\n\x01T\x12\x02\x30\x01"
    node = b"\n\x01r\x12\x04Relu\x1a\x04nope\x2a" + bytes([len(attr)]) + attr
    bad = b"\n" + bytes([len(node)]) + node
    with pytest.raises(client.OpError) as e:
        g.import_graph_def(bad)
    assert "Unknown input node" in e.value.message


@pytest.mark.gpu
def test_real_tensorflow_graph_runs_to_the_known_answer():
    known = json.load(open(GOLDEN_NODES))["known_answer"]
    tf.reset_default_graph()
    g = tf.get_default_graph()
    g.import_graph_def(open(GOLDEN_PB, "rb").read())
    x, y = g.get_tensor_by_name("x:0"), g.get_tensor_by_name("y:0")
    with client.Session(g) as sess:
        sess.run(g.get_operation_by_name("init"))       # NoOp <- ^a/Assign ^b/Assign ^c/Assign
        out = sess.run(y, {x: np.array(known["x"], np.float32).reshape(4, 1)})
        np.testing.assert_array_equal(out, np.array(known["y"], np.float32).reshape(4, 1))
        y2 = sess.run(g.get_tensor_by_name("y2:0"), {x: np.array([[4.0]], np.float32)})
        np.testing.assert_array_equal(y2, [[0.5 * 4 + 3]])   # c = 3 (half_plus_three head)
        # the parser / saver part of the graph is imported but cannot run here
        with pytest.raises(client.OpError) as e:
            sess.run(y)      # x unfed -> needs ParseExample
        assert e.value.error_code == 5 and "ParseExample" in e.value.message
        with pytest.raises(client.OpError):
            sess.run(g.get_operation_by_name("save/restore_all"))
        # the session is still usable afterwards
        out = sess.run(y, {x: np.zeros((2, 1), np.float32)})
        np.testing.assert_array_equal(out, np.full((2, 1), 2.0, np.float32))


# ------------------------------------------------------------------ property test: an independent
# (test-side, pure Python) protobuf encoder writes random GraphDefs; the C++ reader must see
# exactly what was written, and re-serialising must not change what it reads.
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field, payload):
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def _enc_attr(kind, v):
    import struct
    if kind == "s":
        return _ld(2, v)
    if kind == "i":
        return _varint(3 << 3) + _varint(v)
    if kind == "f":
        return _varint(4 << 3 | 5) + struct.pack("<f", v)
    if kind == "b":
        return _varint(5 << 3) + _varint(1 if v else 0)
    if kind == "type":
        return _varint(6 << 3) + _varint(v)
    if kind == "shape":
        return _ld(7, b"".join(_ld(2, (_varint(1 << 3) + _varint(d)) if d else b"") for d in v))
    if kind == "list_i":     # packed or not: both are legal on the wire
        packed, vals = v
        body = _ld(3, b"".join(_varint(x) for x in vals)) if packed else \
            b"".join(_varint(3 << 3) + _varint(x) for x in vals)
        return _ld(1, body)
    if kind == "list_s":
        return _ld(1, b"".join(_ld(2, x) for x in v))
    if kind == "tensor_f32":  # float_val form, possibly fewer values than elements
        shape, vals = v
        tshape = b"".join(_ld(2, _varint(1 << 3) + _varint(d)) for d in shape)
        body = _varint(1 << 3) + _varint(1) + _ld(2, tshape)
        body += _ld(5, b"".join(struct.pack("<f", x) for x in vals))
        return _ld(8, body)
    raise AssertionError(kind)


def _expect_attr(kind, v):
    if kind == "s":
        return "s:" + _quote(v)
    if kind == "i":
        return "i:%d" % v
    if kind == "f":
        return "f:%.9g" % np.float32(v)
    if kind == "b":
        return "b:true" if v else "b:false"
    if kind == "type":
        return "type:%d" % v
    if kind == "shape":
        return "shape:[%s]" % ",".join(map(str, v))
    if kind == "list_i":
        return "list_i:[%s]" % ",".join(map(str, v[1])) if v[1] else None  # empty list: raw
    if kind == "list_s":
        return "list_s:[%s]" % ",".join(_quote(x) for x in v) if v else None
    if kind == "tensor_f32":
        shape, vals = v
        n = int(np.prod(shape)) if shape else 1
        full = ([np.float32(x) for x in vals] + [np.float32(vals[-1])] * n)[:n]
        return "tensor:1:[%s]:%s" % (",".join(map(str, shape)), ",".join("%.9g" % x for x in full[:8]))
    raise AssertionError(kind)


def _quote(b):
    out = '"'
    for c in b:
        if c in (34, 92):
            out += "\\" + chr(c)
        elif c < 32 or c > 126:
            out += "\\%03o" % c
        else:
            out += chr(c)
    return out + '"'


def test_random_graphdefs_read_back_exactly():
    from hypothesis import given, settings, strategies as st
    name = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_/0123456789", min_size=1, max_size=12)
    attr = st.one_of(
        st.tuples(st.just("s"), st.binary(max_size=12)),
        st.tuples(st.just("i"), st.integers(-2**63, 2**63 - 1)),
        st.tuples(st.just("f"), st.floats(width=32, allow_nan=False, allow_infinity=False)),
        st.tuples(st.just("b"), st.booleans()),
        st.tuples(st.just("type"), st.sampled_from([1, 2, 3, 7, 9, 14, 19])),
        st.tuples(st.just("shape"), st.lists(st.integers(0, 10**6), max_size=5)),
        st.tuples(st.just("list_i"), st.tuples(st.booleans(), st.lists(st.integers(-2**40, 2**40), max_size=6))),
        st.tuples(st.just("list_s"), st.lists(st.binary(max_size=6), max_size=4)),
        st.tuples(st.just("tensor_f32"),
                  st.tuples(st.lists(st.integers(1, 4), max_size=3),
                            st.lists(st.floats(width=32, allow_nan=False, allow_infinity=False),
                                     min_size=1, max_size=3))),
    )
    node = st.tuples(name, name, st.lists(name, max_size=3), st.sampled_from(["", "/gpu:0"]),
                     st.dictionaries(st.text(alphabet="abcdefgh_", min_size=1, max_size=6), attr,
                                     max_size=4))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(node, max_size=5))
    def check(nodes):
        blob = b""
        for nm, op, inputs, device, attrs in nodes:
            body = _ld(1, nm.encode()) + _ld(2, op.encode())
            body += b"".join(_ld(3, i.encode()) for i in inputs)
            if device:
                body += _ld(4, device.encode())
            for k, (kind, v) in attrs.items():
                body += _ld(5, _ld(1, k.encode()) + _ld(2, _enc_attr(kind, v)))
            blob += _ld(1, body)
        blob += _ld(4, b"\x08\x15")   # versions { producer: 21 }: carried verbatim
        got = _parse_dump(client.graph_def_to_text(blob)) if nodes else []
        assert len(got) == len(nodes)
        for g, (nm, op, inputs, device, attrs) in zip(got, nodes):
            assert (g["name"], g["op"], g["device"], g["input"]) == (nm, op, device, inputs)
            assert set(g["attr"]) == set(attrs)
            for k, (kind, v) in attrs.items():
                want = _expect_attr(kind, v)
                if want is None:
                    assert g["attr"][k].startswith("raw:")
                else:
                    assert g["attr"][k] == want, (k, kind, v)

    check()
