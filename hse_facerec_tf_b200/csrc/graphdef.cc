// Protobuf wire-format reader for TF GraphDef / NodeDef / AttrValue / TensorProto / TensorShapeProto.
// Field numbers follow tensorflow/core/framework/*.proto (TF 1.x).  Unknown fields are skipped.
#include <cmath>
#include <cstring>

#include "graph.h"

namespace hfr {
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t r = 0;
    int shift = 0;
    while (true) {
      if (p >= end) throw std::runtime_error("graphdef: truncated varint");
      uint8_t b = *p++;
      r |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return r;
      shift += 7;
      if (shift > 63) throw std::runtime_error("graphdef: varint too long");
    }
  }
  Reader sub() {
    uint64_t n = varint();
    if ((uint64_t)(end - p) < n) throw std::runtime_error("graphdef: truncated length-delimited field");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: need(8); p += 8; break;
      case 2: sub(); break;
      case 5: need(4); p += 4; break;
      default: throw std::runtime_error("graphdef: unsupported wire type");
    }
  }
  void need(size_t n) {
    if ((size_t)(end - p) < n) throw std::runtime_error("graphdef: truncated fixed field");
  }
  float f32() {
    need(4);
    float v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  double f64() {
    need(8);
    double v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  std::string str() {
    Reader r = sub();
    return std::string((const char*)r.p, (size_t)(r.end - r.p));
  }
};

void parse_shape(Reader r, std::vector<int64_t>* dims, bool* unknown_rank) {
  dims->clear();
  while (!r.done()) {
    uint64_t key = r.varint();
    int fno = (int)(key >> 3), wt = (int)(key & 7);
    if (fno == 2 && wt == 2) {
      Reader d = r.sub();
      int64_t size = 0;
      while (!d.done()) {
        uint64_t k2 = d.varint();
        if ((k2 >> 3) == 1 && (k2 & 7) == 0)
          size = (int64_t)d.varint();
        else
          d.skip((int)(k2 & 7));
      }
      dims->push_back(size);
    } else if (fno == 3 && wt == 0) {
      if (r.varint() && unknown_rank) *unknown_rank = true;
    } else {
      r.skip(wt);
    }
  }
}

template <typename T>
void append_raw(const uint8_t* p, size_t bytes, std::vector<float>* out) {
  size_t n = bytes / sizeof(T);
  out->reserve(out->size() + n);
  for (size_t i = 0; i < n; ++i) {
    T v;
    memcpy(&v, p + i * sizeof(T), sizeof(T));
    out->push_back((float)v);
  }
}

void parse_tensor(Reader r, HTensor* t) {
  t->dtype = 1;
  t->shape.clear();
  t->f.clear();
  const uint8_t* content = nullptr;
  size_t content_len = 0;
  std::vector<float> vals;
  while (!r.done()) {
    uint64_t key = r.varint();
    int fno = (int)(key >> 3), wt = (int)(key & 7);
    if (fno == 1 && wt == 0) {
      t->dtype = (int)r.varint();
    } else if (fno == 2 && wt == 2) {
      parse_shape(r.sub(), &t->shape, nullptr);
    } else if (fno == 4 && wt == 2) {
      Reader c = r.sub();
      content = c.p;
      content_len = (size_t)(c.end - c.p);
    } else if (fno == 5) {  // float_val, packed or single
      if (wt == 2) {
        Reader c = r.sub();
        while (!c.done()) vals.push_back(c.f32());
      } else if (wt == 5) {
        vals.push_back(r.f32());
      } else {
        r.skip(wt);
      }
    } else if (fno == 6) {  // double_val
      if (wt == 2) {
        Reader c = r.sub();
        while (!c.done()) vals.push_back((float)c.f64());
      } else if (wt == 1) {
        vals.push_back((float)r.f64());
      } else {
        r.skip(wt);
      }
    } else if (fno == 7 || fno == 10 || fno == 11) {  // int_val / int64_val / bool_val
      if (wt == 2) {
        Reader c = r.sub();
        while (!c.done()) vals.push_back((float)(int64_t)c.varint());
      } else if (wt == 0) {
        vals.push_back((float)(int64_t)r.varint());
      } else {
        r.skip(wt);
      }
    } else {
      r.skip(wt);
    }
  }
  int64_t n = t->numel();
  if (content_len) {
    switch (t->dtype) {
      case 1: append_raw<float>(content, content_len, &t->f); break;
      case 2: append_raw<double>(content, content_len, &t->f); break;
      case 3: append_raw<int32_t>(content, content_len, &t->f); break;
      case 4: case 12: append_raw<uint8_t>(content, content_len, &t->f); break;
      case 6: append_raw<int8_t>(content, content_len, &t->f); break;
      case 9: append_raw<int64_t>(content, content_len, &t->f); break;
      case 10: append_raw<uint8_t>(content, content_len, &t->f); break;
      default: throw std::runtime_error("graphdef: unsupported tensor dtype " + std::to_string(t->dtype));
    }
    if ((int64_t)t->f.size() != n) throw std::runtime_error("graphdef: tensor_content size does not match shape");
  } else {
    if (vals.empty()) vals.push_back(0.f);
    t->f = vals;
    // a short *_val list is padded with its last element (TF semantics); one value = broadcast
    if ((int64_t)t->f.size() < n) t->f.resize((size_t)n, vals.back());
  }
}

void parse_attr(Reader r, AttrVal* a) {
  while (!r.done()) {
    uint64_t key = r.varint();
    int fno = (int)(key >> 3), wt = (int)(key & 7);
    if (fno == 2 && wt == 2) {
      a->kind = AttrVal::S;
      a->s = r.str();
    } else if (fno == 3 && wt == 0) {
      a->kind = AttrVal::I;
      a->i = (int64_t)r.varint();
    } else if (fno == 4 && wt == 5) {
      a->kind = AttrVal::F;
      a->f = r.f32();
    } else if (fno == 5 && wt == 0) {
      a->kind = AttrVal::B;
      a->b = r.varint() != 0;
    } else if (fno == 6 && wt == 0) {
      a->kind = AttrVal::TYPE;
      a->i = (int64_t)r.varint();
    } else if (fno == 7 && wt == 2) {
      a->kind = AttrVal::SHAPE;
      parse_shape(r.sub(), &a->shape, &a->unknown_rank);
    } else if (fno == 8 && wt == 2) {
      a->kind = AttrVal::TENSOR;
      parse_tensor(r.sub(), &a->tensor);
    } else if (fno == 1 && wt == 2) {
      a->kind = AttrVal::LIST;
      Reader l = r.sub();
      while (!l.done()) {
        uint64_t k2 = l.varint();
        int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
        if (f2 == 3) {  // list(i), packed or not
          if (w2 == 2) {
            Reader c = l.sub();
            while (!c.done()) a->shape.push_back((int64_t)c.varint());
          } else {
            a->shape.push_back((int64_t)l.varint());
          }
        } else {
          l.skip(w2);
        }
      }
    } else {
      r.skip(wt);
    }
  }
}

void parse_node(Reader r, GNode* n) {
  while (!r.done()) {
    uint64_t key = r.varint();
    int fno = (int)(key >> 3), wt = (int)(key & 7);
    if (fno == 1 && wt == 2) {
      n->name = r.str();
    } else if (fno == 2 && wt == 2) {
      n->op = r.str();
    } else if (fno == 3 && wt == 2) {
      n->inputs.push_back(r.str());
    } else if (fno == 5 && wt == 2) {
      Reader e = r.sub();
      std::string k;
      AttrVal v;
      while (!e.done()) {
        uint64_t k2 = e.varint();
        int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
        if (f2 == 1 && w2 == 2)
          k = e.str();
        else if (f2 == 2 && w2 == 2)
          parse_attr(e.sub(), &v);
        else
          e.skip(w2);
      }
      n->attrs[k] = std::move(v);
    } else {
      r.skip(wt);
    }
  }
}

}  // namespace

void parse_graphdef(const uint8_t* data, size_t size, Graph* g) {
  g->nodes.clear();
  g->index.clear();
  Reader r{data, data + size};
  while (!r.done()) {
    uint64_t key = r.varint();
    int fno = (int)(key >> 3), wt = (int)(key & 7);
    if (fno == 1 && wt == 2) {
      GNode n;
      parse_node(r.sub(), &n);
      g->nodes.push_back(std::move(n));
    } else {
      r.skip(wt);  // library (2), versions (4), ...
    }
  }
  if (g->nodes.empty()) throw std::runtime_error("graphdef: no nodes (not a GraphDef?)");
  for (size_t i = 0; i < g->nodes.size(); ++i) g->index[g->nodes[i].name] = (int)i;
}

}  // namespace hfr
