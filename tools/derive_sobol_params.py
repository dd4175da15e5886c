#!/usr/bin/env python3
"""Recover the Joe-Kuo generating parameters (degree s, polynomial a, initial m_1..m_s) behind the rows of the reference's
direction-number table (src/Util/Sobol.inl: kMatrices[10005][32]) and emit them as a compact C include,
adypt_b200/csrc/sobol_joe_kuo.inc. The table itself is NOT copied: Sobol direction numbers are fully determined by these
published parameters (S. Joe & F. Kuo, "new-joe-kuo-6.21201"), and both the product (adypt_b200/csrc/hostmath.cpp) and the
oracle regenerate the 32 columns from them at start-up; the tests pin the regenerated columns against the reference
(tests/test_oracle_pins.py, via oracle/_ref).

Recurrence (Bratley & Fox): for k > s,
  m_k = 2 a_1 m_{k-1} ^ 4 a_2 m_{k-2} ^ ... ^ 2^{s-1} a_{s-1} m_{k-s+1} ^ 2^s m_{k-s} ^ m_{k-s}
  v_k = m_k << (32 - k)       (column k-1 of the table)
Every m_i is odd, so bit i of (m_{k-i} << i) is set and the bits below are clear: the coefficients a_1, a_2, ... of a row fall
out one after the other from m_{s+1} ^ m_1 ^ (m_1 << s), and the degree is the first s for which all 32 columns reproduce.

Packed form, five 32-bit words per dimension (bit 0 = least significant bit of word 0):
  bits 0..4    s  (0 for dimension 1, the van der Corput column set)
  bits 5..20   a  (s-1 coefficient bits, a_1 the most significant, as Joe & Kuo print it)
  bits 21..    (m_i >> 1) in i-1 bits, i = 1..s   (m_i is odd and below 2^i)
Run where /root/reference exists:  python tools/derive_sobol_params.py > adypt_b200/csrc/sobol_joe_kuo.inc
"""
import re
import sys

REF = "/root/reference/src/Util/Sobol.inl"
MAX_S = 18  # 5 + 16 + s(s-1)/2 bits must fit 160


def read_rows():
    with open(REF) as f:
        txt = f.read()
    body = txt[txt.index("kMatrices"):]
    rows = []
    for m in re.finditer(r"\{([^{}]*)\}", body):
        vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
        if len(vals) == 32:
            rows.append(vals)
    return rows


def columns(s, a, m_init):
    m = list(m_init)
    for k in range(s, 32):
        v = m[k - s] ^ (m[k - s] << s)
        for i in range(1, s):
            if (a >> (s - 1 - i)) & 1:
                v ^= m[k - i] << i
        m.append(v)
    return [(m[k] << (31 - k)) & 0xFFFFFFFF for k in range(32)]


def derive(row):
    if row == [1 << (31 - k) for k in range(32)]:
        return (0, 0, [])  # dimension 1: van der Corput
    m = [row[k] >> (31 - k) for k in range(32)]
    for s in range(1, MAX_S + 1):
        t = m[s] ^ m[0] ^ (m[0] << s)
        a = 0
        for i in range(1, s):
            bit = (t >> i) & 1
            a |= bit << (s - 1 - i)
            if bit:
                t ^= m[s - i] << i
        if columns(s, a, m[:s]) == row:
            return (s, a, m[:s])
    raise RuntimeError("no generating polynomial of degree <= %d" % MAX_S)


def pack(s, a, m):
    bits, pos = s | (a << 5), 21
    assert a < (1 << 16) and s < 32
    for i, mi in enumerate(m, start=1):
        assert mi & 1 and mi < (1 << i)
        bits |= (mi >> 1) << pos
        pos += i - 1
    assert pos <= 160
    return [(bits >> (32 * w)) & 0xFFFFFFFF for w in range(5)]


def main():
    rows = read_rows()
    out = ["// Joe-Kuo Sobol parameters (degree s, polynomial a, initial m_1..m_s) for %d dimensions, five packed 32-bit words" % len(rows),
           "// per dimension; layout and provenance in tools/derive_sobol_params.py, which generated this file.",
           "// bits 0..4 s | bits 5..20 a | then (m_i >> 1) in i-1 bits for i = 1..s"]
    smax = 0
    for row in rows:
        s, a, m = derive(row)
        smax = max(smax, s)
        out.append(", ".join("0x%08xu" % w for w in pack(s, a, m)) + ",")
    out.insert(3, "// highest degree: %d" % smax)
    print("\n".join(out))


if __name__ == "__main__":
    main()
