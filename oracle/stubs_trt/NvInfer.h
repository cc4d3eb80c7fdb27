// Declaration-level stand-in for the few TensorRT types /root/reference/src/SuperPoint.cc, LightGlue.cc and EigenPlaces.cc
// name (TensorRT is not in this image), so that those files can be compiled in place.  The member functions are DEFINED by
// the shim that links them: oracle/ref_nethost_shim.cpp as "no engine" failures (only host-side member functions of the
// wrapper classes are called there), oracle/ref_e2e_shim.cpp functionally (two fixed engines whose enqueueV3 calls back
// into the test, which serves the graphs with the CPU oracle).  Nothing here infers anything.  TEST INFRASTRUCTURE.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace nvinfer1 {
enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorIOMode : int32_t { kNONE = 0, kINPUT = 1, kOUTPUT = 2 };
struct Dims {
  int32_t nbDims = 0;
  int64_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
struct Dims4 : Dims {
  Dims4(int64_t a, int64_t b, int64_t c, int64_t e) {
    nbDims = 4;
    d[0] = a, d[1] = b, d[2] = c, d[3] = e;
  }
};
class ILogger {
 public:
  enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
  virtual void log(Severity severity, const char* msg) noexcept = 0;
  virtual ~ILogger() = default;
};
class IExecutionContext {
 public:
  bool setInputShape(const char* name, const Dims& dims);
  Dims getTensorShape(const char* name) const;
  bool setTensorAddress(const char* name, void* data);
  bool enqueueV3(cudaStream_t stream);
};
class ICudaEngine {
 public:
  IExecutionContext* createExecutionContext();
  int32_t getNbIOTensors() const;
  const char* getIOTensorName(int32_t index) const;
  DataType getTensorDataType(const char* name) const;
  TensorIOMode getTensorIOMode(const char* name) const;
  Dims getTensorShape(const char* name) const;
};
class IRuntime {
 public:
  ICudaEngine* deserializeCudaEngine(const void* blob, std::size_t size);
};
IRuntime* createInferRuntime(ILogger& logger);
}  // namespace nvinfer1
