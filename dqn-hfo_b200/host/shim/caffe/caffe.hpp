// caffe/caffe.hpp — what the reference's dqn.hpp / dqn_main.cpp name from Caffe, without Caffe: the two protobuf
// messages that cross the dqn::DQN interface (caffe_types.hpp), the mode switch (dqn_main.cpp:208-212), the two
// prototxt helpers (dqn_main.cpp:233-246) and opaque Net / Solver / Layer types for the ShareLayer signature
// (dqn.hpp:125).  No Caffe object exists behind them: the nets live in HBM behind a dqnb_handle.
#pragma once
#include "../caffe_types.hpp"
#include <boost/shared_ptr.hpp>
namespace caffe {
class Caffe {
 public:
  enum Brew { CPU, GPU };
  // The hot path always runs on the GPU (libdqn_b200 has no CPU fallback); -gpu=false is accepted and logged.
  static void set_mode(Brew mode);
  static Brew mode();
};
template <typename Dtype> class Layer {};
template <typename Dtype> class Net {};
template <typename Dtype> class Solver {};
// <prefix>_{actor,critic}.prototxt in protobuf text format (host/prototxt.cpp)
void ReadProtoFromTextFileOrDie(const char *filename, NetParameter *proto);
void WriteProtoToTextFile(const NetParameter &proto, const char *filename);
}  // namespace caffe
