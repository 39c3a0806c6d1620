// Flat fp32 weight-blob layout — mirrors styletts-zs_b200/spec.py:weight_entries entry for entry
// (tests/test_abi_cpu.py compares the two through stz_weight_offset).  Host-only, no CUDA.
#pragma once
#include <string>
#include <vector>
#include "../../include/stz.h"

namespace stz {

struct WeightEntry {
  std::string name;
  size_t n;    // elements
  size_t off;  // offset in floats (every entry padded to a multiple of 64 floats)
};

inline std::vector<WeightEntry> build_layout(const stz_config& c, size_t* total) {
  std::vector<WeightEntry> E;
  size_t off = 0;
  auto add = [&](const std::string& name, size_t n) {
    E.push_back({name, n, off});
    off += (n + 63) / 64 * 64;
  };
  const size_t d = c.d_model, Ds = c.d_style, L = c.n_layers;
  const size_t n_mod = (9 * L + 2) * d;
  add("in.w", d * Ds); add("in.b", d);
  add("pos", (size_t)c.n_style * d);
  add("time.w1", d * c.d_time); add("time.b1", d);
  add("time.w2", d * d); add("time.b2", d);
  add("ptext.w", d * c.d_text); add("ptext.b", d);
  add("pprompt.w", d * c.d_prompt); add("pprompt.b", d);
  add("null_pp", d);
  add("ctx_text.w", d * c.d_text); add("ctx_text.b", d);
  add("ctx_prompt.w", d * c.d_prompt); add("ctx_prompt.b", d);
  add("type_emb", 2 * d);
  add("null_tok", d);
  add("mod.w", n_mod * d); add("mod.b", n_mod);
  for (size_t l = 0; l < L; ++l) {
    const std::string p = "l" + std::to_string(l) + ".";
    add(p + "qkv.w", 3 * d * d); add(p + "qkv.b", 3 * d);
    add(p + "o.w", d * d); add(p + "o.b", d);
    add(p + "q2.w", d * d); add(p + "q2.b", d);
    add(p + "kv2.w", 2 * d * d); add(p + "kv2.b", 2 * d);
    add(p + "o2.w", d * d); add(p + "o2.b", d);
    add(p + "ff1.w", (size_t)c.d_ff * d); add(p + "ff1.b", c.d_ff);
    add(p + "ff2.w", d * c.d_ff); add(p + "ff2.b", d);
  }
  add("out.w", Ds * d); add("out.b", Ds);
  const size_t ds = c.d_sty_tok, dh = c.d_hid, h = c.d_hid / 2;
  add("sp.q.w", ds * c.d_text); add("sp.q.b", ds);
  add("sp.k.w", ds * Ds); add("sp.k.b", ds);
  add("sp.v.w", ds * Ds); add("sp.v.b", ds);
  add("sp.o.w", ds * ds); add("sp.o.b", ds);
  for (int l = 0; l < c.n_lstm; ++l) {
    for (const char* dr : {"f", "r"}) {
      const std::string p = "lstm" + std::to_string(l) + "." + dr + ".";
      add(p + "w_ih", 4 * h * (dh + ds)); add(p + "w_hh", 4 * h * h);
      add(p + "b_ih", 4 * h); add(p + "b_hh", 4 * h);
    }
    if (l < c.n_lstm - 1) {
      add("adaln" + std::to_string(l) + ".w", 2 * dh * ds);
      add("adaln" + std::to_string(l) + ".b", 2 * dh);
    }
  }
  add("dur.w", (size_t)c.max_dur * dh); add("dur.b", c.max_dur);
  // prosody (F0 / energy) heads behind the length regulator: appended, so the offsets above never move
  const size_t dp = dh / 2;
  for (const char* dr : {"f", "r"}) {
    const std::string p = std::string("pros.lstm.") + dr + ".";
    add(p + "w_ih", 4 * h * (dh + ds)); add(p + "w_hh", 4 * h * h);
    add(p + "b_ih", 4 * h); add(p + "b_hh", 4 * h);
  }
  add("pros.h1.w", 2 * dp * (dh + ds)); add("pros.h1.b", 2 * dp);
  add("pros.f0.w", dp); add("pros.f0.b", 1);
  add("pros.en.w", dp); add("pros.en.b", 1);
  // guidance-scale embedding of the guidance-conditioned student
  add("gs.w1", d * c.d_time); add("gs.b1", d);
  add("gs.w2", d * d); add("gs.b2", d);
  if (total) *total = off;
  return E;
}

}  // namespace stz
