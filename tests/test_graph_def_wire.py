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
