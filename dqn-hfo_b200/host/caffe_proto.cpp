// caffe_proto.cpp — protobuf wire-format codec for the reference's checkpoint files (see caffe_proto.hpp).
#include "caffe_proto.hpp"

#include <cstdint>
#include <cstring>

namespace caffe_proto {
namespace {

// ---- wire format: tag = (field << 3) | type; type 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32 ----
void put_varint(std::string &o, uint64_t v) {
  while (v >= 0x80) { o.push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
  o.push_back((char)v);
}
void put_tag(std::string &o, int field, int type) { put_varint(o, ((uint64_t)field << 3) | (uint64_t)type); }
void put_bytes(std::string &o, int field, const std::string &s) {
  put_tag(o, field, 2);
  put_varint(o, s.size());
  o.append(s);
}
void put_int(std::string &o, int field, long long v) {
  put_tag(o, field, 0);
  put_varint(o, (uint64_t)v);          // negative int32/int64 are sign-extended to 10 bytes, as protobuf does
}

struct Reader {
  const uint8_t *p, *end;
  bool ok = true;
  Reader(const std::string &s) : p((const uint8_t *)s.data()), end((const uint8_t *)s.data() + s.size()) {}
  Reader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // next field: its number, wire type and (for type 2) payload range; scalars land in `scalar`
  bool next(int *field, int *type, uint64_t *scalar, const uint8_t **b, const uint8_t **e) {
    const uint64_t tag = varint();
    if (!ok) return false;
    *field = (int)(tag >> 3);
    *type = (int)(tag & 7);
    if (*field <= 0) { ok = false; return false; }
    switch (*type) {
      case 0: *scalar = varint(); return ok;
      case 1: if (end - p < 8) { ok = false; return false; } memcpy(scalar, p, 8); p += 8; return true;
      case 5: { if (end - p < 4) { ok = false; return false; } uint32_t v; memcpy(&v, p, 4); *scalar = v; p += 4; return true; }
      case 2: {
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return false; }
        *b = p; *e = p + n; p += n;
        return true;
      }
      default: ok = false; return false;     // groups (3, 4) do not occur in caffe.proto
    }
  }
};

std::string encode_blob(const Blob &b) {
  std::string o;
  {   // shape = 7 { dim = 1, packed int64 }
    std::string dims, shape;
    for (long long d : b.shape) put_varint(dims, (uint64_t)d);
    put_bytes(shape, 1, dims);
    put_bytes(o, 7, shape);
  }
  {   // data = 5, packed float (little-endian, like every host this runs on)
    std::string raw((const char *)b.data.data(), b.data.size() * sizeof(float));
    put_bytes(o, 5, raw);
  }
  return o;
}

bool decode_blob(const uint8_t *b, const uint8_t *e, Blob *out) {
  Reader r(b, e);
  long long legacy[4] = {0, 0, 0, 0};
  bool has_legacy = false;
  while (!r.done()) {
    int f, t; uint64_t s = 0; const uint8_t *pb = nullptr, *pe = nullptr;
    if (!r.next(&f, &t, &s, &pb, &pe)) return false;
    if (f == 7 && t == 2) {                               // BlobShape
      Reader rs(pb, pe);
      while (!rs.done()) {
        int f2, t2; uint64_t s2 = 0; const uint8_t *qb = nullptr, *qe = nullptr;
        if (!rs.next(&f2, &t2, &s2, &qb, &qe)) return false;
        if (f2 == 1 && t2 == 2) { Reader rd(qb, qe); while (!rd.done()) { out->shape.push_back((long long)rd.varint()); if (!rd.ok) return false; } }
        else if (f2 == 1 && t2 == 0) out->shape.push_back((long long)s2);     // unpacked spelling
      }
    } else if (f == 5 && t == 2) {                        // packed floats
      if ((pe - pb) % 4) return false;
      const size_t n = (size_t)(pe - pb) / 4, old = out->data.size();
      out->data.resize(old + n);
      memcpy(out->data.data() + old, pb, n * 4);
    } else if (f == 5 && t == 5) {                        // unpacked float
      const uint32_t u = (uint32_t)s; float v; memcpy(&v, &u, 4); out->data.push_back(v);
    } else if (f >= 1 && f <= 4 && t == 0) {              // legacy num / channels / height / width
      legacy[f - 1] = (long long)s; has_legacy = true;
    }
  }
  if (out->shape.empty() && has_legacy) out->shape.assign(legacy, legacy + 4);
  return true;
}

bool decode_layer(const uint8_t *b, const uint8_t *e, bool v1, Layer *L) {
  const int f_name = v1 ? 4 : 1, f_bottom = v1 ? 2 : 3, f_top = v1 ? 3 : 4, f_blobs = v1 ? 6 : 7;
  Reader r(b, e);
  while (!r.done()) {
    int f, t; uint64_t s = 0; const uint8_t *pb = nullptr, *pe = nullptr;
    if (!r.next(&f, &t, &s, &pb, &pe)) return false;
    if (t != 2) continue;
    if (f == f_name) L->name.assign((const char *)pb, (size_t)(pe - pb));
    else if (!v1 && f == 2) L->type.assign((const char *)pb, (size_t)(pe - pb));
    else if (f == f_bottom) L->bottoms.emplace_back((const char *)pb, (size_t)(pe - pb));
    else if (f == f_top) L->tops.emplace_back((const char *)pb, (size_t)(pe - pb));
    else if (f == f_blobs) { L->blobs.emplace_back(); if (!decode_blob(pb, pe, &L->blobs.back())) return false; }
  }
  return true;
}

}  // namespace

std::string EncodeNet(const Net &net) {
  std::string o;
  put_bytes(o, 1, net.name);
  for (const Layer &L : net.layers) {
    std::string l;
    put_bytes(l, 1, L.name);
    put_bytes(l, 2, L.type);
    for (const std::string &b : L.bottoms) put_bytes(l, 3, b);
    for (const std::string &t : L.tops) put_bytes(l, 4, t);
    for (const Blob &b : L.blobs) put_bytes(l, 7, encode_blob(b));
    put_bytes(o, 100, l);
  }
  return o;
}

bool DecodeNet(const std::string &bytes, Net *net) {
  Reader r(bytes);
  bool any = false;
  while (!r.done()) {
    int f, t; uint64_t s = 0; const uint8_t *pb = nullptr, *pe = nullptr;
    if (!r.next(&f, &t, &s, &pb, &pe)) return false;
    if (f == 1 && t == 2) { net->name.assign((const char *)pb, (size_t)(pe - pb)); any = true; }
    else if ((f == 100 || f == 2) && t == 2) {
      net->layers.emplace_back();
      if (!decode_layer(pb, pe, f == 2, &net->layers.back())) return false;
      any = true;
    }
  }
  return r.ok && any;
}

std::string EncodeSolverState(const SolverState &st) {
  std::string o;
  put_int(o, 1, st.iter);
  put_bytes(o, 2, st.learned_net);
  for (const Blob &b : st.history) put_bytes(o, 3, encode_blob(b));
  put_int(o, 4, st.current_step);
  return o;
}

bool DecodeSolverState(const std::string &bytes, SolverState *st) {
  Reader r(bytes);
  bool has_iter = false;
  while (!r.done()) {
    int f, t; uint64_t s = 0; const uint8_t *pb = nullptr, *pe = nullptr;
    if (!r.next(&f, &t, &s, &pb, &pe)) return false;
    if (f == 1 && t == 0) { st->iter = (int)(int64_t)s; has_iter = true; }
    else if (f == 2 && t == 2) st->learned_net.assign((const char *)pb, (size_t)(pe - pb));
    else if (f == 3 && t == 2) { st->history.emplace_back(); if (!decode_blob(pb, pe, &st->history.back())) return false; }
    else if (f == 4 && t == 0) st->current_step = (int)(int64_t)s;
  }
  return r.ok && has_iter;
}

// ---- mapping to the flat learnable_params order ----------------------------------------------------
std::vector<ParamLayer> ParamLayers(int state_size, const std::vector<int> &hidden, bool critic) {
  std::vector<ParamLayer> v;
  int in = state_size + (critic ? 10 : 0);                 // concat [states | actions(4) | action_params(6)], dqn.cpp:446-448
  for (size_t i = 0; i < hidden.size(); ++i) {
    v.push_back({"ip" + std::to_string(i + 1) + "_layer", hidden[i], in});
    in = hidden[i];
  }
  if (critic) v.push_back({"q_values_layer", 1, in});
  else { v.push_back({"action_layer", 4, in}); v.push_back({"actionpara_layer", 6, in}); }
  return v;
}

long long ParamCount(const std::vector<ParamLayer> &layers) {
  long long n = 0;
  for (const ParamLayer &L : layers) n += (long long)L.out * L.in + L.out;
  return n;
}

Net NetFromFlat(const std::string &net_name, int state_size, const std::vector<int> &hidden, bool critic, const float *flat) {
  Net net;
  net.name = net_name;
  std::string input = critic ? "state_actions" : "states";
  long long off = 0;
  size_t tower = 0;
  for (const ParamLayer &P : ParamLayers(state_size, hidden, critic)) {
    Layer L;
    L.name = P.name;
    L.type = "InnerProduct";
    const bool is_tower = tower < hidden.size();
    const std::string top = is_tower ? "ip" + std::to_string(tower + 1)
                                     : (P.name == "action_layer" ? "actions" : P.name == "actionpara_layer" ? "action_params" : "q_values");
    L.bottoms = {input};
    L.tops = {top};
    Blob w, b;
    w.shape = {P.out, P.in};
    w.data.assign(flat + off, flat + off + (long long)P.out * P.in);
    off += (long long)P.out * P.in;
    b.shape = {P.out};
    b.data.assign(flat + off, flat + off + P.out);
    off += P.out;
    L.blobs = {w, b};
    net.layers.push_back(L);
    if (is_tower) {
      Layer R;
      R.name = "ip" + std::to_string(tower + 1) + "_relu_layer";
      R.type = "ReLU";
      R.bottoms = {top};
      R.tops = {top};
      net.layers.push_back(R);
      input = top;
      ++tower;
    }
  }
  return net;
}

static long long blob_count(const Blob &b) {
  long long n = 1;
  for (long long d : b.shape) n *= d;
  return b.shape.empty() ? (long long)b.data.size() : n;
}

int FlatFromNet(const Net &net, int state_size, const std::vector<int> &hidden, bool critic, float *flat, std::string *err) {
  int copied = 0;
  long long off = 0;
  for (const ParamLayer &P : ParamLayers(state_size, hidden, critic)) {
    const long long nw = (long long)P.out * P.in, nb = P.out;
    for (const Layer &L : net.layers) {
      if (L.name != P.name) continue;
      if (L.blobs.empty()) break;                           // a layer definition without weights (prototxt-like)
      if (L.blobs.size() != 2 || blob_count(L.blobs[0]) != nw || (long long)L.blobs[0].data.size() != nw ||
          blob_count(L.blobs[1]) != nb || (long long)L.blobs[1].data.size() != nb) {
        if (err) *err = "layer " + P.name + ": expected W[" + std::to_string(P.out) + "x" + std::to_string(P.in) + "] and b[" +
                        std::to_string(P.out) + "], file has " + std::to_string(L.blobs.size()) + " blobs of " +
                        std::to_string(L.blobs.empty() ? 0 : (long long)L.blobs[0].data.size()) + " / " +
                        std::to_string(L.blobs.size() < 2 ? 0 : (long long)L.blobs[1].data.size()) + " values";
        return -1;
      }
      memcpy(flat + off, L.blobs[0].data.data(), sizeof(float) * nw);
      memcpy(flat + off + nw, L.blobs[1].data.data(), sizeof(float) * nb);
      ++copied;
      break;
    }
    off += nw + nb;
  }
  return copied;
}

std::vector<Blob> HistoryFromFlat(const std::vector<ParamLayer> &layers, const float *m, const float *v) {
  std::vector<Blob> h;
  for (const float *src : {m, v}) {
    long long off = 0;
    for (const ParamLayer &P : layers) {
      Blob w, b;
      w.shape = {P.out, P.in};
      w.data.assign(src + off, src + off + (long long)P.out * P.in);
      off += (long long)P.out * P.in;
      b.shape = {P.out};
      b.data.assign(src + off, src + off + P.out);
      off += P.out;
      h.push_back(w);
      h.push_back(b);
    }
  }
  return h;
}

bool FlatFromHistory(const std::vector<Blob> &history, const std::vector<ParamLayer> &layers, float *m, float *v, std::string *err) {
  const size_t n = 2 * layers.size();
  if (history.size() != 2 * n) {
    if (err) *err = "solver state has " + std::to_string(history.size()) + " history blobs, an Adam solver over this net has " + std::to_string(2 * n);
    return false;
  }
  for (int half = 0; half < 2; ++half) {
    float *dst = half ? v : m;
    long long off = 0;
    for (size_t i = 0; i < layers.size(); ++i) {
      const long long want[2] = {(long long)layers[i].out * layers[i].in, (long long)layers[i].out};
      for (int k = 0; k < 2; ++k) {
        const Blob &b = history[half * n + 2 * i + k];
        if ((long long)b.data.size() != want[k]) {
          if (err) *err = "history blob " + std::to_string(half * n + 2 * i + k) + " has " + std::to_string(b.data.size()) + " values, expected " + std::to_string(want[k]);
          return false;
        }
        memcpy(dst + off, b.data.data(), sizeof(float) * want[k]);
        off += want[k];
      }
    }
  }
  return true;
}

}  // namespace caffe_proto
