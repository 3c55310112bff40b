// Session / SessionOptions / NewSession -- core/public/session.h, core/public/session_options.h.
#ifndef B200TF_CORE_PUBLIC_SESSION_H_
#define B200TF_CORE_PUBLIC_SESSION_H_

#include <string>
#include <utility>
#include <vector>

#include "tensorflow/core/framework/node_def.h"
#include "tensorflow/core/framework/tensor.h"

namespace tensorflow {

struct SessionOptions {
  std::string target;                       // "" = in-process DirectSession
  int gpu_device_id = 0;                    // which visible GPU hosts this session's graph copy
  size_t gpu_memory_limit_bytes = 0;        // 0 = free - max(300 MiB, 5 %)  (gpu_device.cc:546-561)
  void* collective_comm = nullptr;          // NCCL communicator for B200AllReduce (not owned)
  int num_replicas = 1;
};

// Per-step statistics returned through RunMetadata-like plumbing (step_stats.proto subset).
struct RunStats {
  long long nodes_executed = 0;
  long long kernels_launched = 0;  // libb200tf launch-counter delta over the step
  long long h2d_bytes = 0;
  long long d2h_bytes = 0;
  long long host_enqueue_us = 0;   // host time from Run() entry until every kernel was enqueued
  long long host_total_us = 0;     // host time of the whole Run() including the final device sync
};

class Session {
 public:
  virtual ~Session() {}
  virtual Status Create(const GraphDef& graph) = 0;
  virtual Status Extend(const GraphDef& graph) = 0;
  // session.h: Run(inputs, output_tensor_names, target_node_names, outputs)
  virtual Status Run(const std::vector<std::pair<std::string, Tensor>>& inputs,
                     const std::vector<std::string>& output_tensor_names,
                     const std::vector<std::string>& target_node_names,
                     std::vector<Tensor>* outputs) = 0;
  virtual Status Close() = 0;
  // Additive (the role tf.contrib's StagingArea / prefetch-to-device plays for input pipelines):
  // starts copying a pinned host tensor to the session's device on the host_to_device stream
  // and returns immediately with a device-resident Tensor that Run() accepts as a feed value
  // without another copy -- the copy of step i+1 overlaps the kernels of step i.
  virtual Status StageFeed(const Tensor& host, Tensor* staged) {
    return errors::Unimplemented("StageFeed is not supported by this session");
  }
  virtual const RunStats& last_run_stats() const = 0;
};

Status NewSession(const SessionOptions& options, Session** out_session);

}  // namespace tensorflow
#endif
