// host_selftest.cpp — CPU-only checks of the host mirror (no GPU, no libdqn_b200 calls): flag
// parsing, GetAction semantics, LabelTransitions arithmetic, the in-process HFO stand-in and the
// reward shaping of HFOGameState; with --gpu, the checks that need a DQN object (SelectActions' epsilon branch and the
// order of its random draws).  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "caffe_proto.hpp"
#include "dqn.hpp"
#include "hfo_game.hpp"
#include "shim/flags.hpp"
#include "shim/logging.hpp"

namespace dqn {
DECLARE_double(gamma);
DECLARE_int32(batch_size);
DECLARE_string(hidden);
}
using dqn::FLAGS_batch_size;
using dqn::FLAGS_gamma;
using dqn::FLAGS_hidden;

static int failures = 0;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

// --gpu: the part of the mirror that needs a device.  SelectActions' epsilon branch (dqn.cpp:695-711) and
// GetRandomActorOutput (dqn.cpp:664-682) draw from the DQN's std::mt19937 in a fixed order - one double for the coin flip,
// then per row 4 logits U(-1,1), dash power U(-100,100), three angles U(-180,180), kick power U(0,100), kick angle - and the
// greedy branch consumes exactly the coin flip.  Replayed here on a second engine with the same seed: bit equality.
static void gpu_section() {
  shim::set_flag("seed", "7"); shim::set_flag("batch_size", "32"); shim::set_flag("hidden", "64,32,32,16"); shim::set_flag("memory", "1000");
  const int S = 59;
  caffe::SolverParameter ap, cp;
  ap.set_type("Adam"); cp.set_type("Adam");
  ap.set_base_lr(1e-5f); cp.set_base_lr(1e-3f);
  ap.set_momentum(0.95f); cp.set_momentum(0.95f); ap.set_momentum2(0.999f); cp.set_momentum2(0.999f);
  ap.set_clip_gradients(10.f); cp.set_clip_gradients(10.f);
  ap.net_param_ = dqn::CreateActorNet(S); cp.net_param_ = dqn::CreateCriticNet(S);
  dqn::DQN d(ap, cp, "/tmp/host_selftest_gpu", S, 0);
  std::mt19937 twin(7);
  auto U = [&](float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(twin); };
  std::vector<dqn::InputStates> batch;
  std::mt19937 srng(11);
  for (int i = 0; i < 5; ++i) {
    auto st = std::make_shared<std::vector<float>>(S);
    for (float &v : *st) v = std::uniform_real_distribution<float>(-1.f, 1.f)(srng);
    dqn::InputStates in; in[dqn::kStateInputCount - 1] = st;
    batch.push_back(in);
  }
  // epsilon = 1: the whole batch is random, rows drawn one after the other
  std::vector<dqn::ActorOutput> r = d.SelectActions(batch, 1.0);
  (void)std::uniform_real_distribution<double>(0.0, 1.0)(twin);          // the coin flip
  EXPECT(r.size() == batch.size());
  for (auto &o : r) {
    for (int k = 0; k < 4; ++k) EXPECT(o[k] == U(-1.f, 1.f));
    EXPECT(o[4] == U(-100.f, 100.f)); EXPECT(o[5] == U(-180.f, 180.f)); EXPECT(o[6] == U(-180.f, 180.f));
    EXPECT(o[7] == U(-180.f, 180.f)); EXPECT(o[8] == U(0.f, 100.f)); EXPECT(o[9] == U(-180.f, 180.f));
    EXPECT(o[8] >= 0.f && o[8] <= 100.f && std::fabs(o[4]) <= 100.f);
  }
  // epsilon = 0: greedy, consumes only the coin flip; deterministic; the same rows give the same actions one by one
  std::vector<dqn::ActorOutput> g1 = d.SelectActions(batch, 0.0);
  (void)std::uniform_real_distribution<double>(0.0, 1.0)(twin);
  std::vector<dqn::ActorOutput> g2 = d.SelectActions(batch, 0.0);
  (void)std::uniform_real_distribution<double>(0.0, 1.0)(twin);
  EXPECT(g1 == g2);
  dqn::ActorOutput one = d.SelectAction(batch[3], 0.0);
  (void)std::uniform_real_distribution<double>(0.0, 1.0)(twin);
  for (int k = 0; k < 10; ++k) EXPECT(std::fabs(one[k] - g1[3][k]) <= 1e-6f * (1.f + std::fabs(one[k])));
  // the engines are still in step: the next random output matches draw for draw
  dqn::ActorOutput nxt = d.GetRandomActorOutput();
  for (int k = 0; k < 4; ++k) EXPECT(nxt[k] == U(-1.f, 1.f));
  EXPECT(nxt[4] == U(-100.f, 100.f)); EXPECT(nxt[5] == U(-180.f, 180.f)); EXPECT(nxt[6] == U(-180.f, 180.f));
  EXPECT(nxt[7] == U(-180.f, 180.f)); EXPECT(nxt[8] == U(0.f, 100.f)); EXPECT(nxt[9] == U(-180.f, 180.f));
  // 0 < epsilon < 1: the branch taken is the coin flip's
  for (int t = 0; t < 20; ++t) {
    std::mt19937 peek(twin);                                                   // a copy: look at the coin without consuming it
    const bool expect_random = std::uniform_real_distribution<double>(0.0, 1.0)(peek) < 0.5;
    std::vector<dqn::ActorOutput> m = d.SelectActions(batch, 0.5);
    (void)std::uniform_real_distribution<double>(0.0, 1.0)(twin);
    if (expect_random) { for (auto &o : m) { for (int k = 0; k < 4; ++k) EXPECT(o[k] == U(-1.f, 1.f)); for (int k = 4; k < 10; ++k) (void)U(k == 4 ? -100.f : k == 8 ? 0.f : -180.f, k == 4 || k == 8 ? 100.f : 180.f); } }
    else EXPECT(m == g1);
  }
}

// --write-caffemodel <out.caffemodel> <actor|critic> <state_size> <h1,h2,..> <weights.bin>
// --read-caffemodel  <in.caffemodel>  <actor|critic> <state_size> <h1,h2,..> <weights_out.bin>
// The checkpoint codec on its own (no GPU): flat learnable_params array (float32, Caffe order) <-> NetParameter file.
// tests/test_caffemodel_opencv.py hands the written file to OpenCV's Caffe importer, a third-party reader of the format.
static int caffemodel_tool(int argc, char **argv) {
  if (argc != 7) { std::fprintf(stderr, "usage: %s --write-caffemodel|--read-caffemodel FILE actor|critic S H1,H2,.. WEIGHTS.bin\n", argv[0]); return 2; }
  const bool write = std::string(argv[1]) == "--write-caffemodel", critic = std::string(argv[3]) == "critic";
  const int S = std::atoi(argv[4]);
  std::vector<int> hidden;
  for (const char *c = argv[5]; *c;) { hidden.push_back(std::atoi(c)); while (*c && *c != ',') ++c; if (*c == ',') ++c; }
  const long long n = caffe_proto::ParamCount(caffe_proto::ParamLayers(S, hidden, critic));
  std::vector<float> flat((size_t)n, 0.f);
  auto slurp = [](const char *f, std::string *out) { FILE *fp = std::fopen(f, "rb"); if (!fp) return false; char buf[1 << 16]; size_t k; while ((k = std::fread(buf, 1, sizeof(buf), fp)) > 0) out->append(buf, k); std::fclose(fp); return true; };
  auto dump = [](const char *f, const void *p, size_t bytes) { FILE *fp = std::fopen(f, "wb"); if (!fp) return false; const bool ok = std::fwrite(p, 1, bytes, fp) == bytes; std::fclose(fp); return ok; };
  if (write) {
    std::string raw;
    if (!slurp(argv[6], &raw) || (long long)raw.size() != 4 * n) { std::fprintf(stderr, "weights file: expected %lld floats\n", n); return 1; }
    std::memcpy(flat.data(), raw.data(), raw.size());
    const std::string bytes = caffe_proto::EncodeNet(caffe_proto::NetFromFlat(critic ? "Critic" : "Actor", S, hidden, critic, flat.data()));
    return dump(argv[2], bytes.data(), bytes.size()) ? 0 : 1;
  }
  std::string bytes, err;
  caffe_proto::Net net;
  if (!slurp(argv[2], &bytes) || !caffe_proto::DecodeNet(bytes, &net)) { std::fprintf(stderr, "not a NetParameter file\n"); return 1; }
  const int copied = caffe_proto::FlatFromNet(net, S, hidden, critic, flat.data(), &err);
  if (copied < 0) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
  std::printf("layers copied: %d\n", copied);
  return dump(argv[6], flat.data(), 4 * (size_t)n) ? 0 : 1;
}

// --write-prototxt <out.prototxt> <actor|critic> <state_size> <h1,h2,..> <batch>: the net description dqn_main writes as
// <prefix>_{actor,critic}.prototxt (dqn_main.cpp:232-246), for third-party parsers
static int prototxt_tool(int argc, char **argv) {
  if (argc != 7) { std::fprintf(stderr, "usage: %s --write-prototxt FILE actor|critic S H1,H2,.. BATCH\n", argv[0]); return 2; }
  shim::set_flag("hidden", argv[5]);
  const bool critic = std::string(argv[3]) == "critic";
  const int S = std::atoi(argv[4]);
  dqn::WriteNetPrototxt(critic ? dqn::CreateCriticNet(S) : dqn::CreateActorNet(S), argv[2], std::atoi(argv[6]));
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "--write-prototxt") return prototxt_tool(argc, argv);
  if (argc > 1 && (std::string(argv[1]) == "--write-caffemodel" || std::string(argv[1]) == "--read-caffemodel")) return caffemodel_tool(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--gpu") {
    gpu_section();
    std::printf(failures ? "host_selftest --gpu: %d FAILURES\n" : "host_selftest --gpu: ok\n", failures);
    return failures ? 1 : 0;
  }
  {  // flags
    const char *av[] = {"prog", "-gamma=0.5", "--batch_size", "1024", "-hidden=64,32", "positional"};
    int ac = 6; char **a = const_cast<char **>(av);
    shim::ParseCommandLineFlags(&ac, &a, true);
    EXPECT(FLAGS_gamma == 0.5); EXPECT(FLAGS_batch_size == 1024); EXPECT(FLAGS_hidden == "64,32");
    EXPECT(ac == 2 && std::strcmp(a[1], "positional") == 0);
    shim::set_flag("gamma", "0.99"); shim::set_flag("batch_size", "32"); shim::set_flag("hidden", "1024,512,256,128");
  }
  {  // GetAction (dqn.cpp:196-208): tackle masked, first max wins, param offsets
    dqn::ActorOutput o = {0.1f, 0.5f, 9.0f, 0.5f, 10, 20, 30, 40, 50, 60};
    Action a = dqn::GetAction(o);
    EXPECT(a.action == hfo::TURN && a.arg1 == 30.f && a.arg2 == 0.f);
    o[3] = 0.6f; a = dqn::GetAction(o);
    EXPECT(a.action == hfo::KICK && a.arg1 == 50.f && a.arg2 == 60.f);
    o[0] = 0.7f; a = dqn::GetAction(o);
    EXPECT(a.action == hfo::DASH && a.arg1 == 10.f && a.arg2 == 20.f);
    EXPECT(dqn::PrintActorOutput(o).find("Dash(10.000000, 20.000000)=0.700000") == 0);
  }
  {  // <prefix>_{actor,critic}.prototxt (dqn_main.cpp:232-246): what we write is what we read ...
    shim::set_flag("hidden", "1024,512,256,128");
    for (int critic = 0; critic < 2; ++critic) {
      const caffe::NetParameter np = critic ? dqn::CreateCriticNet(59) : dqn::CreateActorNet(59);
      const std::string text = dqn::NetPrototxt(np, 32);
      EXPECT(text.find("name: \"" + std::string(critic ? "Critic" : "Actor") + "\"") == 0);
      EXPECT(text.find("force_backward: true") != std::string::npos);
      EXPECT(text.find("height: 59") != std::string::npos);
      EXPECT(text.find("name: \"ip4_relu_layer\"") != std::string::npos && text.find("negative_slope: 0.01") != std::string::npos);
      EXPECT((text.find("name: \"q_values_layer\"") != std::string::npos) == (critic == 1));
      EXPECT((text.find("name: \"actionpara_layer\"") != std::string::npos) == (critic == 0));
      EXPECT((text.find("type: \"Concat\"") != std::string::npos) == (critic == 1));
      caffe::NetParameter back;
      dqn::ParseNetPrototxtOrDie(text, "roundtrip", critic != 0, &back);
      EXPECT(back.hidden == np.hidden && back.state_size == 59 && back.critic == (critic != 0) && back.name() == np.name());
    }
    // ... and a hand-edited file of the same family (comments, other widths and depth, single-line messages,
    // the `field: { }` spelling) gives the tower the user asked for
    const char *edited =
        "# a narrower actor\n"
        "name: \"Actor\"  force_backward: true\n"
        "layer { name: \"state_input_layer\" type: \"MemoryData\" top: \"states\" top: \"dummy1\"\n"
        "        memory_data_param { batch_size: 32 channels: 1 height: 77 width: 1 } }\n"
        "layer { name: \"silence\" type: \"Silence\" bottom: \"dummy1\" }\n"
        "layer { name: \"ip1_layer\" type: \"InnerProduct\" bottom: \"states\" top: \"ip1\"\n"
        "        inner_product_param: { num_output: 300 weight_filler { type: \"gaussian\" std: 0.01 } } }\n"
        "layer { name: \"ip1_relu_layer\" type: \"ReLU\" bottom: \"ip1\" top: \"ip1\" relu_param { negative_slope: 0.01 } }\n"
        "layer { name: \"ip2_layer\" type: \"InnerProduct\" bottom: \"ip1\" top: \"ip2\" inner_product_param { num_output: 200 } }\n"
        "layer { name: \"ip2_relu_layer\" type: \"ReLU\" bottom: \"ip2\" top: \"ip2\" relu_param { negative_slope: 1e-2 } }\n"
        "layer { name: \"action_layer\" type: \"InnerProduct\" bottom: \"ip2\" top: \"actions\" inner_product_param { num_output: 4 } }\n"
        "layer { name: \"actionpara_layer\" type: \"InnerProduct\" bottom: \"ip2\" top: \"action_params\" inner_product_param { num_output: 6 } }\n";
    caffe::NetParameter ed;
    dqn::ParseNetPrototxtOrDie(edited, "edited", false, &ed);
    EXPECT(ed.hidden.size() == 2 && ed.hidden[0] == 300 && ed.hidden[1] == 200 && ed.state_size == 77 && !ed.critic);
  }
  {  // Caffe checkpoint files (caffe_proto.cpp): NetParameter / SolverState in protobuf wire format
    namespace cp = caffe_proto;
    const std::vector<int> hidden = {5, 3};
    const int S = 4;
    for (int critic = 0; critic < 2; ++critic) {
      const std::vector<cp::ParamLayer> pl = cp::ParamLayers(S, hidden, critic != 0);
      const long long n = cp::ParamCount(pl);
      EXPECT(n == (critic ? (14 * 5 + 5) + (5 * 3 + 3) + (3 + 1) : (4 * 5 + 5) + (5 * 3 + 3) + (3 * 4 + 4) + (3 * 6 + 6)));
      std::vector<float> w((size_t)n), m((size_t)n), v((size_t)n);
      for (long long i = 0; i < n; ++i) { w[i] = 0.25f * (float)i - 3.f; m[i] = 1e-3f * (float)i; v[i] = 1e-6f * (float)(i * i); }
      // weights: flat -> NetParameter bytes -> flat
      const std::string bytes = cp::EncodeNet(cp::NetFromFlat(critic ? "Critic" : "Actor", S, hidden, critic != 0, w.data()));
      cp::Net net;
      EXPECT(cp::DecodeNet(bytes, &net));
      EXPECT(net.name == (critic ? "Critic" : "Actor") && net.layers.size() == pl.size() + hidden.size());
      EXPECT(net.layers[0].name == "ip1_layer" && net.layers[0].type == "InnerProduct" && net.layers[1].type == "ReLU");
      EXPECT(net.layers[0].blobs.size() == 2 && net.layers[0].blobs[0].shape.size() == 2 && net.layers[0].blobs[0].shape[0] == 5);
      std::vector<float> back((size_t)n, -7.f);
      std::string err;
      EXPECT(cp::FlatFromNet(net, S, hidden, critic != 0, back.data(), &err) == (int)pl.size());
      EXPECT(back == w);
      // CopyTrainedLayersFrom semantics: a layer missing from the file keeps its values, a foreign layer is ignored
      cp::Net partial = net;
      partial.layers.erase(partial.layers.begin());            // drop ip1_layer
      cp::Layer foreign; foreign.name = "conv1"; foreign.type = "Convolution"; foreign.blobs.resize(1); foreign.blobs[0].data = {1.f, 2.f};
      partial.layers.push_back(foreign);
      std::vector<float> kept((size_t)n, 9.f);
      EXPECT(cp::FlatFromNet(partial, S, hidden, critic != 0, kept.data(), &err) == (int)pl.size() - 1);
      const long long first = (long long)pl[0].out * pl[0].in + pl[0].out;
      bool ok = true;
      for (long long i = 0; i < n; ++i) ok = ok && kept[i] == (i < first ? 9.f : w[i]);
      EXPECT(ok);
      // a known layer with the wrong shape is an error, not a silent skip
      cp::Net bad = net;
      bad.layers[0].blobs[0].data.pop_back();
      EXPECT(cp::FlatFromNet(bad, S, hidden, critic != 0, kept.data(), &err) == -1 && err.find("ip1_layer") != std::string::npos);
      // solver state: iter + Adam history (m for every blob, then v), learned_net path
      cp::SolverState st;
      st.iter = 123456; st.learned_net = "/tmp/x_actor_iter_123456.caffemodel"; st.history = cp::HistoryFromFlat(pl, m.data(), v.data());
      cp::SolverState st2;
      EXPECT(cp::DecodeSolverState(cp::EncodeSolverState(st), &st2));
      EXPECT(st2.iter == 123456 && st2.learned_net == st.learned_net && st2.history.size() == 4 * pl.size() && st2.current_step == 0);
      std::vector<float> m2((size_t)n), v2((size_t)n);
      EXPECT(cp::FlatFromHistory(st2.history, pl, m2.data(), v2.data(), &err));
      EXPECT(m2 == m && v2 == v);
      st2.history.pop_back();
      EXPECT(!cp::FlatFromHistory(st2.history, pl, m2.data(), v2.data(), &err));
    }
    // bytes assembled by hand from the protobuf encoding rules (independent of the encoder above):
    // NetParameter{ name: "N", layer{ name: "q_values_layer", blobs{ shape{dim:[1,2]} data:[1.5,-2] }, blobs{ shape{dim:[1]} data:[0.5] } } }
    // plus an unknown varint field (5: force_backward) and an unknown fixed32 field that must be skipped
    auto f32 = [](float x) { std::string s(4, '\0'); std::memcpy(&s[0], &x, 4); return s; };
    const std::string blob_w = std::string("\x3a\x04\x0a\x02\x01\x02", 6) + std::string("\x2a\x08", 2) + f32(1.5f) + f32(-2.f);
    const std::string blob_b = std::string("\x3a\x03\x0a\x01\x01", 5) + std::string("\x2a\x04", 2) + f32(0.5f);
    std::string layer = std::string("\x0a\x0e", 2) + "q_values_layer";
    layer += std::string("\x3a", 1) + std::string(1, (char)blob_w.size()) + blob_w;
    layer += std::string("\x3a", 1) + std::string(1, (char)blob_b.size()) + blob_b;
    std::string hand = std::string("\x0a\x01N", 3) + std::string("\x28\x01", 2) + std::string("\x7d", 1) + f32(3.f);
    hand += std::string("\xa2\x06", 2) + std::string(1, (char)layer.size()) + layer;
    cp::Net hn;
    EXPECT(cp::DecodeNet(hand, &hn));
    EXPECT(hn.name == "N" && hn.layers.size() == 1 && hn.layers[0].name == "q_values_layer" && hn.layers[0].blobs.size() == 2);
    EXPECT(hn.layers[0].blobs[0].shape == (std::vector<long long>{1, 2}) && hn.layers[0].blobs[0].data == (std::vector<float>{1.5f, -2.f}));
    EXPECT(hn.layers[0].blobs[1].data == (std::vector<float>{0.5f}));
    // the same layer in the pre-2015 V1 spelling: NetParameter.layers = 2, V1LayerParameter{ name = 4, blobs = 6 }, legacy dims
    const std::string v1blob = std::string("\x08\x01\x10\x01\x18\x01\x20\x02", 8) + std::string("\x2a\x08", 2) + f32(1.5f) + f32(-2.f);
    std::string v1layer = std::string("\x22\x0e", 2) + "q_values_layer" + std::string("\x32", 1) + std::string(1, (char)v1blob.size()) + v1blob;
    const std::string v1net = std::string("\x12", 1) + std::string(1, (char)v1layer.size()) + v1layer;
    cp::Net vn;
    EXPECT(cp::DecodeNet(v1net, &vn));
    EXPECT(vn.layers.size() == 1 && vn.layers[0].name == "q_values_layer" && vn.layers[0].blobs.size() == 1);
    EXPECT(vn.layers[0].blobs[0].shape == (std::vector<long long>{1, 1, 1, 2}) && vn.layers[0].blobs[0].data.size() == 2);
    // garbage is rejected, not mis-read
    cp::Net gn;
    EXPECT(!cp::DecodeNet(std::string("DQNBW001\x00\x00\x00\x00", 12), &gn) || gn.layers.empty());
    EXPECT(!cp::DecodeNet(std::string("\xa2\x06\x7f", 3), &gn));     // layer length beyond the end of the buffer
  }
  {  // NumStateFeatures (hfo_game.hpp:14-16)
    EXPECT(NumStateFeatures(1) == 59 && NumStateFeatures(3) == 77);
  }
  {  // in-process environment: deterministic, bounded, ends episodes; reward shaping is finite
    for (int rep = 0; rep < 2; ++rep) {
      hfo::HFOEnvironment env;
      env.configure(59, 50, 7);
      static double first_total = 0;
      double total = 0;
      int episodes = 0, steps = 0;
      std::mt19937 rng(3);
      for (int e = 0; e < 5; ++e) {
        HFOGameState game(env.getUnum());
        env.act(hfo::DASH, 0, 0);
        game.update(env);
        EXPECT(!game.episode_over);
        while (!game.episode_over) {
          const std::vector<float> &f = env.getState();
          EXPECT((int)f.size() == 59);
          for (float v : f) EXPECT(std::isfinite(v) && v >= -1.0001f && v <= 1.0001f);
          Action a = GetRandomHFOAction(rng);
          env.act(a.action, a.arg1, a.arg2);
          game.update(env);
          const float r = game.reward();
          EXPECT(std::isfinite(r));
          ++steps;
        }
        EXPECT(game.steps <= 51);
        total += game.total_reward;
        ++episodes;
      }
      EXPECT(episodes == 5 && steps > 5);
      if (rep == 0) first_total = total; else EXPECT(total == first_total);
    }
  }
  {  // GetRandomHFOAction ranges (hfo_game.cpp:80-107)
    std::mt19937 rng(1);
    for (int i = 0; i < 1000; ++i) {
      Action a = GetRandomHFOAction(rng);
      EXPECT(a.action >= hfo::DASH && a.action <= hfo::KICK);
      if (a.action == hfo::KICK) EXPECT(a.arg1 >= 0.f && a.arg1 <= 100.f);
      EXPECT(std::fabs(a.arg2) <= 180.f);
    }
  }
  std::printf(failures ? "host_selftest: %d FAILURES\n" : "host_selftest: ok\n", failures);
  return failures ? 1 : 0;
}
