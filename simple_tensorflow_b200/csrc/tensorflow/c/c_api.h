/* c_api.h -- the graph/session subset of the reference's tensorflow/c/c_api.h that a front-end
 * needs to build a graph of the hot-path ops and run it (same names, argument meaning and
 * error behaviour: c_api.h:134-153 status, :197-240 tensors, :270-291 session options,
 * :301-530 graph construction, :774 TF_GraphOperationByName, :932-1022 sessions).
 * Entry points that need protobufs (TF_GraphToGraphDef, TF_ImportGraphDef, TF_SetConfig...) are
 * not provided: there is no protoc in this environment.  B200TF_* functions are additive.
 */
#ifndef B200TF_C_C_API_H_
#define B200TF_C_C_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_CAPI_EXPORT __attribute__((visibility("default")))

TF_CAPI_EXPORT extern const char* TF_Version(void);

typedef enum {
  TF_FLOAT = 1, TF_DOUBLE = 2, TF_INT32 = 3, TF_UINT8 = 4, TF_INT16 = 5, TF_INT8 = 6,
  TF_STRING = 7, TF_COMPLEX64 = 8, TF_INT64 = 9, TF_BOOL = 10, TF_BFLOAT16 = 14, TF_HALF = 19
} TF_DataType;
TF_CAPI_EXPORT extern size_t TF_DataTypeSize(TF_DataType dt);

typedef enum {
  TF_OK = 0, TF_CANCELLED = 1, TF_UNKNOWN = 2, TF_INVALID_ARGUMENT = 3, TF_DEADLINE_EXCEEDED = 4,
  TF_NOT_FOUND = 5, TF_ALREADY_EXISTS = 6, TF_PERMISSION_DENIED = 7, TF_UNAUTHENTICATED = 16,
  TF_RESOURCE_EXHAUSTED = 8, TF_FAILED_PRECONDITION = 9, TF_ABORTED = 10, TF_OUT_OF_RANGE = 11,
  TF_UNIMPLEMENTED = 12, TF_INTERNAL = 13, TF_UNAVAILABLE = 14, TF_DATA_LOSS = 15
} TF_Code;

typedef struct TF_Status TF_Status;
TF_CAPI_EXPORT extern TF_Status* TF_NewStatus(void);
TF_CAPI_EXPORT extern void TF_DeleteStatus(TF_Status*);
TF_CAPI_EXPORT extern void TF_SetStatus(TF_Status* s, TF_Code code, const char* msg);
TF_CAPI_EXPORT extern TF_Code TF_GetCode(const TF_Status* s);
TF_CAPI_EXPORT extern const char* TF_Message(const TF_Status* s);

typedef struct TF_Tensor TF_Tensor;
TF_CAPI_EXPORT extern TF_Tensor* TF_NewTensor(TF_DataType, const int64_t* dims, int num_dims,
                                              void* data, size_t len,
                                              void (*deallocator)(void* data, size_t len,
                                                                  void* arg),
                                              void* deallocator_arg);
/* Allocates page-locked host memory when a GPU is present (the reference's GPU-compatible CPU
 * device does the same for tensors that feed a GPU, gpu_device_factory.cc:69-107). */
TF_CAPI_EXPORT extern TF_Tensor* TF_AllocateTensor(TF_DataType, const int64_t* dims, int num_dims,
                                                   size_t len);
TF_CAPI_EXPORT extern void TF_DeleteTensor(TF_Tensor*);
TF_CAPI_EXPORT extern TF_DataType TF_TensorType(const TF_Tensor*);
TF_CAPI_EXPORT extern int TF_NumDims(const TF_Tensor*);
TF_CAPI_EXPORT extern int64_t TF_Dim(const TF_Tensor* tensor, int dim_index);
TF_CAPI_EXPORT extern size_t TF_TensorByteSize(const TF_Tensor*);
TF_CAPI_EXPORT extern void* TF_TensorData(const TF_Tensor*);

typedef struct TF_SessionOptions TF_SessionOptions;
TF_CAPI_EXPORT extern TF_SessionOptions* TF_NewSessionOptions(void);
TF_CAPI_EXPORT extern void TF_SetTarget(TF_SessionOptions* options, const char* target);
TF_CAPI_EXPORT extern void TF_DeleteSessionOptions(TF_SessionOptions*);
/* Additive (stand-ins for ConfigProto fields that would need protobuf): */
TF_CAPI_EXPORT extern void B200TF_SetGpuDevice(TF_SessionOptions* options, int gpu_id);
TF_CAPI_EXPORT extern void B200TF_SetGpuMemoryLimit(TF_SessionOptions* options, size_t bytes);
/* comm: a communicator from b200_nccl_comm_init_rank (include/b200_ops.h); not owned. */
TF_CAPI_EXPORT extern void B200TF_SetCollective(TF_SessionOptions* options, void* comm,
                                                int num_replicas);

typedef struct TF_Graph TF_Graph;
TF_CAPI_EXPORT extern TF_Graph* TF_NewGraph(void);
TF_CAPI_EXPORT extern void TF_DeleteGraph(TF_Graph*);

typedef struct TF_OperationDescription TF_OperationDescription;
typedef struct TF_Operation TF_Operation;
typedef struct TF_Input { TF_Operation* oper; int index; } TF_Input;
typedef struct TF_Output { TF_Operation* oper; int index; } TF_Output;

TF_CAPI_EXPORT extern TF_OperationDescription* TF_NewOperation(TF_Graph* graph,
                                                               const char* op_type,
                                                               const char* oper_name);
TF_CAPI_EXPORT extern void TF_SetDevice(TF_OperationDescription* desc, const char* device);
TF_CAPI_EXPORT extern void TF_AddInput(TF_OperationDescription* desc, TF_Output input);
TF_CAPI_EXPORT extern void TF_AddInputList(TF_OperationDescription* desc, const TF_Output* inputs,
                                           int num_inputs);
TF_CAPI_EXPORT extern void TF_AddControlInput(TF_OperationDescription* desc, TF_Operation* input);
TF_CAPI_EXPORT extern void TF_SetAttrString(TF_OperationDescription* desc, const char* attr_name,
                                            const void* value, size_t length);
TF_CAPI_EXPORT extern void TF_SetAttrInt(TF_OperationDescription* desc, const char* attr_name,
                                         int64_t value);
TF_CAPI_EXPORT extern void TF_SetAttrIntList(TF_OperationDescription* desc, const char* attr_name,
                                             const int64_t* values, int num_values);
TF_CAPI_EXPORT extern void TF_SetAttrFloat(TF_OperationDescription* desc, const char* attr_name,
                                           float value);
TF_CAPI_EXPORT extern void TF_SetAttrBool(TF_OperationDescription* desc, const char* attr_name,
                                          unsigned char value);
TF_CAPI_EXPORT extern void TF_SetAttrType(TF_OperationDescription* desc, const char* attr_name,
                                          TF_DataType value);
TF_CAPI_EXPORT extern void TF_SetAttrShape(TF_OperationDescription* desc, const char* attr_name,
                                           const int64_t* dims, int num_dims);
TF_CAPI_EXPORT extern void TF_SetAttrTensor(TF_OperationDescription* desc, const char* attr_name,
                                            TF_Tensor* value, TF_Status* status);
/* On failure returns NULL and the description is consumed either way (c_api.h:505-515). */
TF_CAPI_EXPORT extern TF_Operation* TF_FinishOperation(TF_OperationDescription* desc,
                                                       TF_Status* status);
TF_CAPI_EXPORT extern const char* TF_OperationName(TF_Operation* oper);
TF_CAPI_EXPORT extern const char* TF_OperationOpType(TF_Operation* oper);
TF_CAPI_EXPORT extern int TF_OperationNumOutputs(TF_Operation* oper);
TF_CAPI_EXPORT extern TF_DataType TF_OperationOutputType(TF_Output oper_out);
TF_CAPI_EXPORT extern int TF_OperationNumInputs(TF_Operation* oper);
TF_CAPI_EXPORT extern TF_Operation* TF_GraphOperationByName(TF_Graph* graph, const char* oper_name);
/* ---- inspecting operations (c_api.h:536-563, 585-760): what an importer needs to walk a graph */
TF_CAPI_EXPORT extern TF_Output TF_OperationInput(TF_Input oper_in);
TF_CAPI_EXPORT extern int TF_OperationNumControlInputs(TF_Operation* oper);
TF_CAPI_EXPORT extern int TF_OperationGetControlInputs(TF_Operation* oper,
                                                       TF_Operation** control_inputs,
                                                       int max_control_inputs);
typedef enum TF_AttrType {
  TF_ATTR_STRING = 0, TF_ATTR_INT = 1, TF_ATTR_FLOAT = 2, TF_ATTR_BOOL = 3, TF_ATTR_TYPE = 4,
  TF_ATTR_SHAPE = 5, TF_ATTR_TENSOR = 6, TF_ATTR_PLACEHOLDER = 7, TF_ATTR_FUNC = 8
} TF_AttrType;
typedef struct TF_AttrMetadata {
  unsigned char is_list;
  int64_t list_size;
  TF_AttrType type;
  int64_t total_size; /* string: bytes; shape: number of dims (-1 unknown rank); c_api.h:608-626 */
} TF_AttrMetadata;
/* Attributes this runtime carries as opaque bytes (list(shape), list(float), func, string tensors)
 * report TF_ATTR_PLACEHOLDER with total_size = their serialized size and cannot be read back. */
TF_CAPI_EXPORT extern TF_AttrMetadata TF_OperationGetAttrMetadata(TF_Operation* oper,
                                                                  const char* attr_name,
                                                                  TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrString(TF_Operation* oper, const char* attr_name,
                                                     void* value, size_t max_length,
                                                     TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrInt(TF_Operation* oper, const char* attr_name,
                                                  int64_t* value, TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrIntList(TF_Operation* oper, const char* attr_name,
                                                      int64_t* values, int max_values,
                                                      TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrFloat(TF_Operation* oper, const char* attr_name,
                                                    float* value, TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrBool(TF_Operation* oper, const char* attr_name,
                                                   unsigned char* value, TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrType(TF_Operation* oper, const char* attr_name,
                                                   TF_DataType* value, TF_Status* status);
TF_CAPI_EXPORT extern void TF_OperationGetAttrShape(TF_Operation* oper, const char* attr_name,
                                                    int64_t* value, int num_dims,
                                                    TF_Status* status);
/* Returns a new host tensor holding a copy of the attribute (caller deletes it). */
TF_CAPI_EXPORT extern void TF_OperationGetAttrTensor(TF_Operation* oper, const char* attr_name,
                                                     TF_Tensor** value, TF_Status* status);
/* Additive: the names of all attributes of an operation, one per line (malloc'ed; free() it). */
TF_CAPI_EXPORT extern char* B200TF_OperationAttrNames(TF_Operation* oper);
/* Iterate the operations of a graph: start with *pos = 0; returns NULL at the end (c_api.h:763). */
TF_CAPI_EXPORT extern TF_Operation* TF_GraphNextOperation(TF_Graph* graph, size_t* pos);

/* ---- GraphDef import / export in protobuf wire format (c_api.h:160-180, 786-836).
 * The codec is hand-written (core/framework/graph_def_wire.cc): there is no protoc here. */
typedef struct TF_Buffer {
  const void* data;
  size_t length;
  void (*data_deallocator)(void* data, size_t length);
} TF_Buffer;
TF_CAPI_EXPORT extern TF_Buffer* TF_NewBufferFromString(const void* proto, size_t proto_len);
TF_CAPI_EXPORT extern TF_Buffer* TF_NewBuffer(void);
TF_CAPI_EXPORT extern void TF_DeleteBuffer(TF_Buffer*);
TF_CAPI_EXPORT extern TF_Buffer TF_GetBuffer(TF_Buffer* buffer);
/* Writes the graph as a serialized GraphDef (versions.producer = 21, the reference's
 * TF_GRAPH_DEF_VERSION, core/public/version.h:90). */
TF_CAPI_EXPORT extern void TF_GraphToGraphDef(TF_Graph* graph, TF_Buffer* output_graph_def,
                                              TF_Status* status);
typedef struct TF_ImportGraphDefOptions TF_ImportGraphDefOptions;
TF_CAPI_EXPORT extern TF_ImportGraphDefOptions* TF_NewImportGraphDefOptions(void);
TF_CAPI_EXPORT extern void TF_DeleteImportGraphDefOptions(TF_ImportGraphDefOptions* opts);
TF_CAPI_EXPORT extern void TF_ImportGraphDefOptionsSetPrefix(TF_ImportGraphDefOptions* opts,
                                                             const char* prefix);
/* Imports every node of a serialized GraphDef.  Departure from the reference: a node whose op
 * type is not registered here (savers, string ops, ...) is imported as an opaque node instead of
 * failing the import; a Session::Run whose pruned sub-graph contains such a node fails with
 * NOT_FOUND "Op type not registered".  `options` may be NULL. */
TF_CAPI_EXPORT extern void TF_GraphImportGraphDef(TF_Graph* graph, const TF_Buffer* graph_def,
                                                  const TF_ImportGraphDefOptions* options,
                                                  TF_Status* status);
/* Additive: a line-per-node text dump of a serialized GraphDef (malloc'ed; free() it). */
TF_CAPI_EXPORT extern char* B200TF_GraphDefToText(const void* proto, size_t proto_len,
                                                  TF_Status* status);

typedef struct TF_Session TF_Session;
TF_CAPI_EXPORT extern TF_Session* TF_NewSession(TF_Graph* graph, const TF_SessionOptions* opts,
                                                TF_Status* status);
TF_CAPI_EXPORT extern void TF_CloseSession(TF_Session*, TF_Status* status);
TF_CAPI_EXPORT extern void TF_DeleteSession(TF_Session*, TF_Status* status);
/* run_options / run_metadata must be NULL (protobuf buffers).  Input tensors are HOST buffers;
 * output tensors are newly allocated host tensors owned by the caller. */
TF_CAPI_EXPORT extern void TF_SessionRun(TF_Session* session, const void* run_options,
                                         const TF_Output* inputs, TF_Tensor* const* input_values,
                                         int ninputs, const TF_Output* outputs,
                                         TF_Tensor** output_values, int noutputs,
                                         const TF_Operation* const* target_opers, int ntargets,
                                         void* run_metadata, TF_Status*);

/* Additive introspection: step statistics of the last TF_SessionRun on this session
 * (what StepStats/RunMetadata would carry). */
typedef struct B200TF_RunStats {
  int64_t nodes_executed, kernels_launched, h2d_bytes, d2h_bytes, host_enqueue_us, host_total_us;
} B200TF_RunStats;
TF_CAPI_EXPORT extern void B200TF_SessionLastRunStats(TF_Session*, B200TF_RunStats* out);
/* The CUstream every kernel of this session is enqueued on (for CUDA-event timing). */
TF_CAPI_EXPORT extern void* B200TF_SessionStream(TF_Session*);
/* Input staging (prefetch to device): starts the host->device copy of a pinned host tensor on
 * the session's host_to_device stream and returns at once.  The returned TF_Tensor is
 * DEVICE-RESIDENT (TF_TensorData is a device pointer: do not dereference it on the host); pass it
 * as an input value of TF_SessionRun, which then uses it without another copy, ordered behind the
 * staging copy.  Staging step i+1 before running step i overlaps the copy with the kernels.
 * The host tensor may be deleted right away (the session keeps its buffer until the copy has
 * been consumed).  Delete the staged tensor with TF_DeleteTensor after the Run that used it,
 * and in any case BEFORE TF_DeleteSession: its memory belongs to the session's device arena. */
TF_CAPI_EXPORT extern TF_Tensor* B200TF_SessionStageTensor(TF_Session*, TF_Tensor* host,
                                                           TF_Status* status);
/* Registered ops / kernels ("Op:DEVICE:label"), newline-separated; caller frees with free(). */
TF_CAPI_EXPORT extern char* B200TF_ListRegisteredOps(void);
TF_CAPI_EXPORT extern char* B200TF_ListRegisteredKernels(void);
/* TF_LoadLibrary (c_api.h:1116): dlopen()s a kernel library so its static REGISTER_OP /
 * REGISTER_KERNEL_BUILDER initialisers run (framework/load_library.cc:46-110). */
typedef struct TF_Library TF_Library;
TF_CAPI_EXPORT extern TF_Library* TF_LoadLibrary(const char* library_filename, TF_Status* status);
TF_CAPI_EXPORT extern void TF_DeleteLibraryHandle(TF_Library* lib_handle);

#ifdef __cplusplus
}
#endif
#endif
