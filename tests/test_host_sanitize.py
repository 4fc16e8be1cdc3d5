"""advanced-hpc-lbm_b200/host/lbm_io.c under AddressSanitizer + UBSan: the obstacle parser fed
damaged files (the reference's parser, d2q9-bgk.c:2826-2857, dies with a message on each of
these; ours must do the same and never touch memory out of bounds), and the "%.12E"
formatter against printf on random floats."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "advanced-hpc-lbm_b200", "host")

DRIVER = r"""
#include "lbm_io.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
  if (argc == 2) {                       /* formatter: N random finite floats vs printf */
    unsigned long long s = 88172645463325252ULL;
    char a[64], b[64];
    long bad = 0, n = atol(argv[1]);
    for (long i = 0; i < n; i++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      union { unsigned u; float f; } v;
      v.u = (unsigned)(s >> 16);
      if (!isfinite(v.f)) continue;
      a[lbm_format_e12(a, (double)v.f)] = 0;
      snprintf(b, sizeof b, "%.12E", (double)v.f);
      bad += strcmp(a, b) != 0;
    }
    printf("%ld\n", bad);
    return bad != 0;
  }
  int nx = atoi(argv[2]), ny = atoi(argv[3]);
  uint32_t* bits = lbm_read_obstacle_bits(argv[1], nx, ny);
  long blocked = 0;
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) blocked += lbm_obstacle_bit(bits, nx, i, j);
  printf("%ld\n", blocked);
  free(bits);
  return 0;
}
"""


@pytest.fixture(scope="module")
def sanitized(tmp_path_factory):
    d = tmp_path_factory.mktemp("asan")
    src = d / "driver.c"
    src.write_text(DRIVER)
    exe = str(d / "driver")
    cmd = ["gcc", "-std=c99", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fopenmp", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, str(src), os.path.join(HOST, "lbm_io.c"),
           "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("gcc cannot link the sanitizer runtimes here: " + r.stderr[-200:])
    return exe, d


def _clean(r):
    return "Sanitizer" not in r.stderr and "runtime error" not in r.stderr


def test_formatter_matches_printf_under_sanitizers(sanitized):
    exe, _ = sanitized
    r = subprocess.run([exe, "300000"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0" and _clean(r), r.stderr[-400:]


def test_obstacle_parser_survives_damaged_files(sanitized):
    exe, d = sanitized
    rows = ["%d %d 1\n" % (x, y) for y in (0, 63) for x in range(64)] + ["%d %d 1\n" % (0, y) for y in range(1, 63)]
    base = "".join(rows)
    r = subprocess.run([exe, _write(d, base), "64", "64"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == str(len(set(rows))) and _clean(r)
    rng = random.Random(20240229)
    junk = ["99999999999999999999", "-1", " ", "\n\n", "1 2", "\r", "\0", "64 0 1\n", "0 64 1\n", "3 3 2\n", "x"]
    died = 0
    for _ in range(80):
        s = list(base[:rng.randint(0, len(base))])
        for _ in range(rng.randint(1, 6)):
            if not s:
                break
            i = rng.randrange(len(s))
            op = rng.random()
            if op < 0.3:
                s[i] = rng.choice("0123456789 -\n\tx.e+")
            elif op < 0.6:
                del s[i]
            else:
                s.insert(i, rng.choice(junk))
        r = subprocess.run([exe, _write(d, "".join(s)), str(rng.choice([64, 1, 33])), str(rng.choice([64, 2]))],
                           capture_output=True, text=True)
        assert _clean(r) and r.returncode in (0, 1), (r.returncode, r.stderr[-400:])
        if r.returncode == 1:
            assert "Error at line" in r.stderr       # the reference's die() format, d2q9-bgk.c:3001-3007
            died += 1
    assert died > 10                                  # the damage was real


def _write(d, text):
    p = d / "obstacles.dat"
    p.write_text(text)
    return str(p)
