// GraphDef <-> protobuf wire format, without protoc.
//
// Reads and writes the binary encoding of the reference's graph.proto / node_def.proto /
// attr_value.proto / tensor.proto / tensor_shape.proto (field numbers cited next to each parser in
// graph_def_wire.cc), so graphs serialized by a real TensorFlow 1.0 front-end can be replayed by
// this runtime and graphs built here can be handed to one.  Everything this runtime does not model
// (list(shape), list(float), functions, string tensors, the function library, versions) is kept
// as raw bytes and written back unchanged.
#ifndef B200TF_CORE_FRAMEWORK_GRAPH_DEF_WIRE_H_
#define B200TF_CORE_FRAMEWORK_GRAPH_DEF_WIRE_H_

#include <string>

#include "tensorflow/core/framework/node_def.h"

namespace tensorflow {

Status ParseGraphDef(const void* data, size_t size, GraphDef* out);
void SerializeGraphDef(const GraphDef& graph, std::string* out);
// One line per node: "node\t<name>\t<op>\t<device>\t<inputs,>\t<attr=summary;...>" (sorted attrs).
std::string GraphDefDebugString(const GraphDef& graph);

}  // namespace tensorflow
#endif
