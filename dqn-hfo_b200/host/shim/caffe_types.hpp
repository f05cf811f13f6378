// caffe_types.hpp (included through <caffe/caffe.hpp>) — plain-struct stand-ins for the two Caffe protobuf messages that appear in the
// reference's public signatures (dqn.hpp:58-60, :204-205).  Only the fields dqn_main.cpp:229-262
// fills are present; accessor names follow protobuf's generated API so caller code reads the same.
#pragma once
#include <string>
#include <vector>

namespace caffe {

struct NetParameter {
  std::string name_;
  bool force_backward_ = true;
  int state_size = 0;
  bool critic = false;
  std::vector<int> hidden;           // tower widths (dqn.cpp:425,:449)
  const std::string &name() const { return name_; }
  void set_name(const std::string &n) { name_ = n; }
  void set_force_backward(bool b) { force_backward_ = b; }
  void CopyFrom(const NetParameter &other) { *this = other; }
};

struct SolverParameter {
  std::string type_ = "Adam";        // dqn_main.cpp:30
  float momentum_ = 0.95f;           // :31
  float momentum2_ = 0.999f;         // :32
  float base_lr_ = 1e-5f;            // :33-34
  float clip_gradients_ = 10.f;      // :35
  std::string lr_policy_ = "fixed";  // :36
  int max_iter_ = 10000000;          // :37
  float delta_ = 1e-8f;              // Caffe default
  std::string snapshot_prefix_;
  std::string net_;
  NetParameter net_param_;
  void set_type(const std::string &t) { type_ = t; }
  void set_momentum(float v) { momentum_ = v; }
  void set_momentum2(float v) { momentum2_ = v; }
  void set_base_lr(float v) { base_lr_ = v; }
  void set_clip_gradients(float v) { clip_gradients_ = v; }
  void set_lr_policy(const std::string &p) { lr_policy_ = p; }
  void set_max_iter(int v) { max_iter_ = v; }
  void set_delta(float v) { delta_ = v; }
  void set_snapshot_prefix(const std::string &p) { snapshot_prefix_ = p; }
  void set_net(const std::string &n) { net_ = n; }
  NetParameter *mutable_net_param() { return &net_param_; }
  const std::string &type() const { return type_; }
  float momentum() const { return momentum_; }
  float momentum2() const { return momentum2_; }
  float base_lr() const { return base_lr_; }
  float clip_gradients() const { return clip_gradients_; }
  float delta() const { return delta_; }
  int max_iter() const { return max_iter_; }
  const std::string &snapshot_prefix() const { return snapshot_prefix_; }
};

}  // namespace caffe
