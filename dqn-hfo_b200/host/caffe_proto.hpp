// caffe_proto.hpp — the two Caffe protobuf files the reference's checkpoints consist of
// (`<prefix>_{actor,critic}_iter_N.caffemodel` = NetParameter, `.solverstate` = SolverState; written by
// Solver::Snapshot, dqn.cpp:589-590, read by Solver::Restore / Net::CopyTrainedLayersFrom, dqn.cpp:525-557),
// encoded / decoded directly in protobuf wire format: this image has no protobuf library, and only a
// handful of fields matter (layer names and their weight blobs, the solver iteration, the Adam history).
//
// Field numbers are those of BVLC Caffe's src/caffe/proto/caffe.proto at the commit the reference pins
// (README.md:7-10): NetParameter{name=1, layers(V1)=2, layer=100}, LayerParameter{name=1, type=2,
// bottom=3, top=4, blobs=7}, V1LayerParameter{bottom=2, top=3, name=4, blobs=6},
// BlobProto{num=1, channels=2, height=3, width=4, data=5, diff=6, shape=7}, BlobShape{dim=1},
// SolverState{iter=1, learned_net=2, history=3, current_step=4}.  Unknown fields are skipped by wire type.
#pragma once
#include <string>
#include <vector>

namespace caffe_proto {

struct Blob {
  std::vector<long long> shape;
  std::vector<float> data;
};
struct Layer {
  std::string name, type;
  std::vector<std::string> bottoms, tops;
  std::vector<Blob> blobs;
};
struct Net {
  std::string name;
  std::vector<Layer> layers;
};
struct SolverState {
  int iter = 0;
  std::string learned_net;
  std::vector<Blob> history;     // AdamSolver: first N blobs = m, next N = v, N = learnable blobs of the net
  int current_step = 0;
};

std::string EncodeNet(const Net &net);
bool DecodeNet(const std::string &bytes, Net *net);                  // false: not a well-formed NetParameter
std::string EncodeSolverState(const SolverState &st);
bool DecodeSolverState(const std::string &bytes, SolverState *st);

// ---- the two nets of this path as Caffe sees them (dqn.cpp:400-454) --------------------------------
// Parametrised layers in Net::learnable_params order: ip1_layer .. ipN_layer, then action_layer +
// actionpara_layer (actor) or q_values_layer (critic); each owns W [out x in] and b [out].
struct ParamLayer { std::string name; int out, in; };
std::vector<ParamLayer> ParamLayers(int state_size, const std::vector<int> &hidden, bool critic);
long long ParamCount(const std::vector<ParamLayer> &layers);
// flat learnable_params array (the order of dqnb_get/set_params) <-> NetParameter
Net NetFromFlat(const std::string &net_name, int state_size, const std::vector<int> &hidden, bool critic, const float *flat);
// Copies every layer of `net` whose name and blob shapes match into `flat` (Net::CopyTrainedLayersFrom
// semantics: source layers the target does not have are ignored, target layers missing from the source
// keep their values).  Returns the number of layers copied; a known layer with the wrong shape is an error
// (returns -1 and describes it in *err).
int FlatFromNet(const Net &net, int state_size, const std::vector<int> &hidden, bool critic, float *flat, std::string *err);
// Adam history (m then v, each in learnable_params order) <-> SolverState.history
std::vector<Blob> HistoryFromFlat(const std::vector<ParamLayer> &layers, const float *m, const float *v);
bool FlatFromHistory(const std::vector<Blob> &history, const std::vector<ParamLayer> &layers, float *m, float *v, std::string *err);

}  // namespace caffe_proto
