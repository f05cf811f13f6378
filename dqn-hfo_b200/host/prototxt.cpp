// prototxt.cpp — the `<prefix>_actor.prototxt` / `<prefix>_critic.prototxt` files of the reference
// (dqn_main.cpp:232-246): written in protobuf text format when absent, read back when present, so a
// user can change the tower of a run by editing the file exactly as upstream.
//
// Upstream these files are full caffe::NetParameter messages.  This build has no protobuf, so the
// writer emits the text that WriteProtoToTextFile produces for the nets built by CreateActorNet /
// CreateCriticNet (dqn.cpp:400-454: MemoryData inputs, Silence, [Concat], the ip<i>_layer /
// ip<i>_relu_layer tower, the head InnerProduct layers, [EuclideanLoss]) and the reader is a small
// text-format parser that accepts any file of that family: it extracts what the kernels are
// parametrised by (tower widths, state size) and refuses, loudly, whatever they do not implement
// (other layer types inside the tower, a leaky slope other than 0.01, heads of the wrong width).
#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "dqn.hpp"
#include "shim/flags.hpp"
#include "shim/logging.hpp"

namespace dqn {
namespace {

// ---- a protobuf text-format message: ordered (field, scalar | sub-message) pairs ------------------
struct Msg;
struct Field {
  std::string name, scalar;          // scalar: string contents, number or identifier as written
  std::shared_ptr<Msg> msg;          // set for `name { ... }`
};
struct Msg {
  std::vector<Field> fields;
  std::vector<const Field *> all(const std::string &n) const {
    std::vector<const Field *> v;
    for (const Field &f : fields) if (f.name == n) v.push_back(&f);
    return v;
  }
  const Field *first(const std::string &n) const {
    for (const Field &f : fields) if (f.name == n) return &f;
    return nullptr;
  }
  std::string str(const std::string &n, const std::string &dflt = "") const {
    const Field *f = first(n);
    return f && !f->msg ? f->scalar : dflt;
  }
  const Msg *sub(const std::string &n) const {
    const Field *f = first(n);
    return f && f->msg ? f->msg.get() : nullptr;
  }
};

class Parser {
 public:
  Parser(const std::string &text, const std::string &file) : t_(text), file_(file) {}
  Msg parse() {
    Msg m = message(false);
    skip();
    CHECK(pos_ == t_.size()) << file_ << ":" << line_ << ": unexpected '}'";
    return m;
  }

 private:
  void skip() {
    while (pos_ < t_.size()) {
      const char c = t_[pos_];
      if (c == '#') { while (pos_ < t_.size() && t_[pos_] != '\n') ++pos_; }
      else if (c == '\n') { ++line_; ++pos_; }
      else if (std::isspace((unsigned char)c) || c == ',' || c == ';') ++pos_;
      else break;
    }
  }
  std::string ident() {
    const size_t b = pos_;
    while (pos_ < t_.size() && (std::isalnum((unsigned char)t_[pos_]) || t_[pos_] == '_' || t_[pos_] == '.' || t_[pos_] == '-' || t_[pos_] == '+')) ++pos_;
    CHECK(pos_ > b) << file_ << ":" << line_ << ": expected a field name or value, found '" << t_[b] << "'";
    return t_.substr(b, pos_ - b);
  }
  std::string value() {
    skip();
    CHECK(pos_ < t_.size()) << file_ << ": value expected at end of file";
    if (t_[pos_] == '"' || t_[pos_] == '\'') {
      const char q = t_[pos_++];
      std::string s;
      while (pos_ < t_.size() && t_[pos_] != q) {
        if (t_[pos_] == '\\' && pos_ + 1 < t_.size()) ++pos_;
        s.push_back(t_[pos_++]);
      }
      CHECK(pos_ < t_.size()) << file_ << ":" << line_ << ": unterminated string";
      ++pos_;
      return s;
    }
    return ident();
  }
  Msg message(bool nested) {
    Msg m;
    for (;;) {
      skip();
      if (pos_ >= t_.size()) { CHECK(!nested) << file_ << ": missing '}' at end of file"; return m; }
      if (t_[pos_] == '}') { if (nested) ++pos_; return m; }
      Field f;
      f.name = ident();
      skip();
      if (pos_ < t_.size() && t_[pos_] == ':') {
        ++pos_;
        skip();
        if (pos_ < t_.size() && t_[pos_] == '{') { ++pos_; f.msg = std::make_shared<Msg>(message(true)); }
        else f.scalar = value();
      } else {
        CHECK(pos_ < t_.size() && t_[pos_] == '{') << file_ << ":" << line_ << ": ':' or '{' expected after " << f.name;
        ++pos_;
        f.msg = std::make_shared<Msg>(message(true));
      }
      m.fields.push_back(std::move(f));
    }
  }
  const std::string &t_;
  std::string file_;
  size_t pos_ = 0;
  int line_ = 1;
};

// ---- writer helpers --------------------------------------------------------------------------------
struct Out {
  std::ostringstream o;
  int depth = 0;
  void line(const std::string &s) { o << std::string(2 * depth, ' ') << s << "\n"; }
  void open(const std::string &s) { line(s + " {"); ++depth; }
  void close() { --depth; line("}"); }
  void layer(const std::string &name, const std::string &type, const std::vector<std::string> &bottoms,
             const std::vector<std::string> &tops) {
    open("layer");
    line("name: \"" + name + "\"");
    line("type: \"" + type + "\"");
    for (const std::string &b : bottoms) line("bottom: \"" + b + "\"");
    for (const std::string &t : tops) line("top: \"" + t + "\"");
  }
  void memory_data(const std::string &name, const std::vector<std::string> &tops, int n, int c, int h, int w) {
    layer(name, "MemoryData", {}, tops);
    open("memory_data_param");
    line("batch_size: " + std::to_string(n));
    line("channels: " + std::to_string(c));
    line("height: " + std::to_string(h));
    line("width: " + std::to_string(w));
    close();
    close();
  }
  void inner_product(const std::string &name, const std::string &bottom, const std::string &top, int num_output) {
    layer(name, "InnerProduct", {bottom}, {top});
    open("inner_product_param");
    line("num_output: " + std::to_string(num_output));
    open("weight_filler");
    line("type: \"gaussian\"");
    line("std: 0.01");
    close();
    close();
    close();
  }
  void relu(const std::string &name, const std::string &blob) {
    layer(name, "ReLU", {blob}, {blob});
    open("relu_param");
    line("negative_slope: 0.01");
    close();
    close();
  }
};

bool is_head(const std::string &layer_name) {
  return layer_name == "action_layer" || layer_name == "actionpara_layer" || layer_name == q_values_layer_name;
}

}  // namespace

std::string NetPrototxt(const caffe::NetParameter &np, int batch_size) {
  Out w;
  w.line("name: \"" + np.name() + "\"");
  w.line(std::string("force_backward: ") + (np.force_backward_ ? "true" : "false"));
  w.memory_data(state_input_layer_name, {states_blob_name, "dummy1"}, batch_size, kStateInputCount, np.state_size, 1);
  std::string input = states_blob_name;
  if (np.critic) {
    w.memory_data(action_input_layer_name, {actions_blob_name, "dummy2"}, batch_size, kStateInputCount, kActionSize, 1);
    w.memory_data(action_params_input_layer_name, {action_params_blob_name, "dummy3"}, batch_size, kStateInputCount, kActionParamSize, 1);
    w.memory_data(target_input_layer_name, {targets_blob_name, "dummy4"}, batch_size, 1, 1, 1);
    w.layer("silence", "Silence", {"dummy1", "dummy2", "dummy3", "dummy4"}, {});
    w.close();
    w.layer("concat", "Concat", {states_blob_name, actions_blob_name, action_params_blob_name}, {"state_actions"});
    w.open("concat_param");
    w.line("axis: 2");
    w.close();
    w.close();
    input = "state_actions";
  } else {
    w.layer("silence", "Silence", {"dummy1"}, {});
    w.close();
  }
  for (size_t i = 0; i < np.hidden.size(); ++i) {     // Tower (dqn.cpp:400-416)
    const std::string top = "ip" + std::to_string(i + 1);
    w.inner_product(top + "_layer", input, top, np.hidden[i]);
    w.relu(top + "_relu_layer", top);
    input = top;
  }
  if (np.critic) {
    w.inner_product(q_values_layer_name, input, q_values_blob_name, 1);
    w.layer("loss", "EuclideanLoss", {q_values_blob_name, targets_blob_name}, {loss_blob_name});
    w.close();
  } else {
    w.inner_product("action_layer", input, actions_blob_name, kActionSize);
    w.inner_product("actionpara_layer", input, action_params_blob_name, kActionParamSize);
  }
  return w.o.str();
}

void WriteNetPrototxt(const caffe::NetParameter &np, const std::string &filename, int batch_size) {
  std::ofstream f(filename);
  CHECK(f.good()) << "cannot write " << filename;
  f << NetPrototxt(np, batch_size);
}

void ParseNetPrototxtOrDie(const std::string &text, const std::string &origin, bool critic, caffe::NetParameter *np) {
  Parser parser(text, origin);
  const Msg net = parser.parse();
  np->set_name(net.str("name", critic ? "Critic" : "Actor"));
  np->set_force_backward(net.str("force_backward", "true") == "true");
  np->critic = critic;
  np->hidden.clear();
  np->state_size = 0;
  std::vector<const Field *> layers = net.all("layer");
  if (layers.empty()) layers = net.all("layers");          // pre-2015 Caffe spelling
  CHECK(!layers.empty()) << origin << ": no layer in the net";
  std::map<std::string, int> heads;
  std::string chain_blob;                                   // top of the tower so far
  bool expect_relu = false;
  for (const Field *lf : layers) {
    CHECK(lf->msg) << origin << ": 'layer' must be a message";
    const Msg &L = *lf->msg;
    const std::string name = L.str("name"), type = L.str("type");
    if (type == "MemoryData") {
      if (name == state_input_layer_name) {
        const Msg *p = L.sub("memory_data_param");
        CHECK(p) << origin << ": " << name << " has no memory_data_param";
        CHECK(p->str("channels", "1") == "1" && p->str("width", "1") == "1")
            << origin << ": states are [batch, " << kStateInputCount << ", state_size, 1] (dqn.cpp:422-423)";
        np->state_size = std::atoi(p->str("height", "0").c_str());
      }
    } else if (type == "InnerProduct") {
      const Msg *p = L.sub("inner_product_param");
      CHECK(p && p->first("num_output")) << origin << ": " << name << " has no inner_product_param.num_output";
      const int n = std::atoi(p->str("num_output").c_str());
      CHECK_GT(n, 0) << origin << ": " << name << ": bad num_output";
      CHECK(p->str("bias_term", "true") == "true") << origin << ": " << name << ": bias_term: false is not implemented";
      if (is_head(name)) {
        heads[name] = n;
      } else {
        CHECK(!expect_relu) << origin << ": tower layer before " << name << " has no ReLU (dqn.cpp:407-409)";
        CHECK(heads.empty()) << origin << ": tower layer " << name << " after a head layer";
        np->hidden.push_back(n);
        chain_blob = L.str("top");
        expect_relu = true;
      }
    } else if (type == "ReLU") {
      CHECK(expect_relu && L.str("bottom") == chain_blob && L.str("top") == chain_blob)
          << origin << ": " << name << ": only in-place ReLU layers directly after a tower InnerProduct are implemented";
      const Msg *p = L.sub("relu_param");
      const double slope = p ? std::atof(p->str("negative_slope", "0").c_str()) : 0.0;
      CHECK(std::fabs(slope - 0.01) < 1e-9) << origin << ": " << name << ": negative_slope " << slope
                                              << " (the kernels implement the reference's 0.01, dqn.cpp:300)";
      expect_relu = false;
    } else if (type == "Silence" || type == "Concat" || type == "EuclideanLoss") {
      // fixed plumbing of the two nets (dqn.cpp:424, :446-448, :451-452)
    } else {
      LOG(FATAL) << origin << ": layer " << name << " of type " << type << " is not implemented by this build "
                 << "(InnerProduct + leaky ReLU towers only, SURVEY 8f-4)";
    }
  }
  CHECK(!expect_relu) << origin << ": last tower layer has no ReLU";
  CHECK(!np->hidden.empty()) << origin << ": no tower (ip<i>_layer) found";
  CHECK_GT(np->state_size, 0) << origin << ": no " << state_input_layer_name;
  if (critic) {
    CHECK(heads.count(q_values_layer_name) && heads[q_values_layer_name] == 1) << origin << ": " << q_values_layer_name << " must have num_output 1";
  } else {
    CHECK(heads.count("action_layer") && heads["action_layer"] == kActionSize) << origin << ": action_layer must have num_output " << kActionSize;
    CHECK(heads.count("actionpara_layer") && heads["actionpara_layer"] == kActionParamSize)
        << origin << ": actionpara_layer must have num_output " << kActionParamSize;
  }
}

void ReadNetPrototxtOrDie(const std::string &filename, bool critic, caffe::NetParameter *np) {
  std::ifstream f(filename);
  CHECK(f.good()) << "Failed to open " << filename;
  std::stringstream ss;
  ss << f.rdbuf();
  ParseNetPrototxtOrDie(ss.str(), filename, critic, np);
}

}  // namespace dqn

// The two Caffe helpers dqn_main.cpp:233-246 calls on <prefix>_{actor,critic}.prototxt.  Which of the two nets a
// file describes is read off its layers (only the critic has a q_values_layer).
namespace dqn { DECLARE_int32(batch_size); }
namespace caffe {
void ReadProtoFromTextFileOrDie(const char *filename, NetParameter *proto) {
  std::ifstream f(filename);
  CHECK(f.good()) << "Failed to open " << filename;
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  const bool critic = text.find(std::string("\"") + dqn::q_values_layer_name + "\"") != std::string::npos;
  dqn::ParseNetPrototxtOrDie(text, filename, critic, proto);
}
void WriteProtoToTextFile(const NetParameter &proto, const char *filename) {
  dqn::WriteNetPrototxt(proto, filename, dqn::FLAGS_batch_size);
}
}  // namespace caffe
