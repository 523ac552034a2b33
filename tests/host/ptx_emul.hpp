// ptx_emul.hpp -- a tiny interpreter for the PTX integer carry-chain instructions the field
// arithmetic of pbc_b200/csrc/fp.cuh and fq_small.cuh is written in.
//
// TEST INFRASTRUCTURE.  tests/host/make_host_fp.py rewrites every `asm volatile("..." : outs : ins)`
// statement of those headers into ptx(<string>, {&outs...}, {ins...}); this file executes the
// string with the PTX ISA semantics (one condition-code carry flag CC.CF, written by the .cc forms,
// read by addc / subc / madc), so the very template code the kernels compile -- every N, FULL or
// not -- runs on the CPU and is compared with big-integer arithmetic.
//   add{c}{.cc}.u32  d, a, b        d = a + b (+ CF)            .cc: CF = carry out
//   sub{c}{.cc}.u32  d, a, b        d = a - b (- CF)            .cc: CF = borrow out
//   mul.{lo,hi}.u32  d, a, b
//   mad{c}.{lo,hi}{.cc}.u32 d,a,b,c d = {lo,hi}(a b) + c (+ CF) .cc: CF = carry out
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

namespace ptxemu {

struct Operand { bool imm; uint32_t val; int idx; };
struct Insn { int op; bool carry_in, cc, hi; int nsrc; Operand dst, src[3]; };
enum { ADD, SUB, MUL, MAD };

inline uint32_t& CF() { static thread_local uint32_t cf = 0; return cf; }
inline uint64_t& executed() { static uint64_t n = 0; return n; }
// 32x32 products issued: every mul.lo / mad{c}.lo stands for one IMAD.WIDE.U32 on the device (its .hi
// partner is the upper half of the same product)
inline uint64_t& products() { static uint64_t n = 0; return n; }

inline Operand parse_operand(const std::string& t) {
  Operand o{};
  if (t[0] == '%') { o.imm = false; o.idx = atoi(t.c_str() + 1); }
  else { o.imm = true; o.val = (uint32_t)strtoul(t.c_str(), nullptr, 0); }
  return o;
}

inline std::vector<Insn> parse(const char* text) {
  std::vector<Insn> prog;
  std::string s(text);
  size_t pos = 0;
  while (pos < s.size()) {
    size_t end = s.find(';', pos);
    if (end == std::string::npos) end = s.size();
    std::string st = s.substr(pos, end - pos);
    pos = end + 1;
    size_t b = st.find_first_not_of(" \t\n");
    if (b == std::string::npos) continue;
    st = st.substr(b);
    size_t sp = st.find(' ');
    std::string opc = st.substr(0, sp), rest = st.substr(sp + 1);
    Insn in{};
    std::vector<std::string> parts;
    size_t p = 0;
    while (p <= opc.size()) {
      size_t d = opc.find('.', p);
      if (d == std::string::npos) d = opc.size();
      parts.push_back(opc.substr(p, d - p));
      p = d + 1;
    }
    const std::string& base = parts[0];
    if (base == "add" || base == "addc") { in.op = ADD; in.carry_in = base == "addc"; in.nsrc = 2; }
    else if (base == "sub" || base == "subc") { in.op = SUB; in.carry_in = base == "subc"; in.nsrc = 2; }
    else if (base == "mul") { in.op = MUL; in.nsrc = 2; }
    else if (base == "mad" || base == "madc") { in.op = MAD; in.carry_in = base == "madc"; in.nsrc = 3; }
    else { fprintf(stderr, "ptx_emul: unknown opcode `%s'\n", opc.c_str()); abort(); }
    bool typed = false;
    for (size_t i = 1; i < parts.size(); i++) {
      if (parts[i] == "cc") in.cc = true;
      else if (parts[i] == "hi") in.hi = true;
      else if (parts[i] == "lo") in.hi = false;
      else if (parts[i] == "u32") typed = true;
      else { fprintf(stderr, "ptx_emul: unknown modifier in `%s'\n", opc.c_str()); abort(); }
    }
    if (!typed) { fprintf(stderr, "ptx_emul: `%s' is not .u32\n", opc.c_str()); abort(); }
    std::vector<std::string> ops;
    p = 0;
    while (p < rest.size()) {
      size_t c = rest.find(',', p);
      if (c == std::string::npos) c = rest.size();
      std::string t = rest.substr(p, c - p);
      size_t tb = t.find_first_not_of(" \t"), te = t.find_last_not_of(" \t");
      ops.push_back(t.substr(tb, te - tb + 1));
      p = c + 1;
    }
    if ((int)ops.size() != in.nsrc + 1) { fprintf(stderr, "ptx_emul: operand count in `%s'\n", st.c_str()); abort(); }
    in.dst = parse_operand(ops[0]);
    if (in.dst.imm) { fprintf(stderr, "ptx_emul: immediate destination\n"); abort(); }
    for (int i = 0; i < in.nsrc; i++) in.src[i] = parse_operand(ops[i + 1]);
    prog.push_back(in);
  }
  return prog;
}

// operands are numbered outputs first, then inputs (GCC extended-asm rule); "+r" outputs are read too
inline void run(const std::vector<Insn>& prog, std::initializer_list<uint32_t*> outs, std::initializer_list<uint32_t> ins) {
  uint32_t* o[8];
  uint32_t iv[8];
  int no = 0, ni = 0;
  for (uint32_t* p : outs) o[no++] = p;
  for (uint32_t v : ins) iv[ni++] = v;
  auto rd = [&](const Operand& x) -> uint32_t {
    if (x.imm) return x.val;
    if (x.idx < no) return *o[x.idx];
    if (x.idx - no >= ni) { fprintf(stderr, "ptx_emul: operand %%%d out of range\n", x.idx); abort(); }
    return iv[x.idx - no];
  };
  uint32_t cf = CF();
  for (const Insn& in : prog) {
    uint64_t r;
    uint32_t cin = in.carry_in ? cf : 0;
    switch (in.op) {
      case ADD: r = (uint64_t)rd(in.src[0]) + rd(in.src[1]) + cin; if (in.cc) cf = (uint32_t)(r >> 32); break;
      case SUB: r = (uint64_t)rd(in.src[0]) - rd(in.src[1]) - cin; if (in.cc) cf = (uint32_t)((r >> 32) & 1); break;
      case MUL: { uint64_t m = (uint64_t)rd(in.src[0]) * rd(in.src[1]); r = in.hi ? m >> 32 : (uint32_t)m; if (!in.hi) __atomic_fetch_add(&products(), 1, __ATOMIC_RELAXED); break; }
      default: {
        uint64_t m = (uint64_t)rd(in.src[0]) * rd(in.src[1]);
        r = (uint64_t)(in.hi ? (uint32_t)(m >> 32) : (uint32_t)m) + rd(in.src[2]) + cin;
        if (in.cc) cf = (uint32_t)(r >> 32);
        if (!in.hi) __atomic_fetch_add(&products(), 1, __ATOMIC_RELAXED);
      }
    }
    if (in.dst.idx >= no) { fprintf(stderr, "ptx_emul: write to an input operand\n"); abort(); }
    *o[in.dst.idx] = (uint32_t)r;
  }
  CF() = cf;
  executed() += prog.size();
}
// one parsed program per asm statement of the source (SITE numbers them): no lookup on the hot path
template <int SITE>
inline void ptx_at(const char* text, std::initializer_list<uint32_t*> outs, std::initializer_list<uint32_t> ins) {
  static const std::vector<Insn> prog = parse(text);
  run(prog, outs, ins);
}
inline void ptx(const char* text, std::initializer_list<uint32_t*> outs, std::initializer_list<uint32_t> ins) {
  run(parse(text), outs, ins);
}

}  // namespace ptxemu
