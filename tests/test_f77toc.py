"""The FORTRAN 77 -> C translator (oracle/f77toc.py) on small programs with known answers: the constructs the
parity argument leans on -- default-real literals, implicit typing, mixed-mode arithmetic, integer division, DO trip
counts, terminal-label `continue`, computed array bounds, by-reference arguments, `save`."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SRC = """
      subroutine t1(x,y,n,out)
c     mixed precision: 0.01*x multiplies by the REAL*4 literal; s is implicit real*4; k integer division
      implicit double precision (a-h,o-z)
      real*4 s
      dimension out(10)
      s = 0.1
      out(1) = 0.01*x
      out(2) = 0.01d0*x
      out(3) = s*3
      out(4) = 7/2
      out(5) = 7/2.
      out(6) = x**2 - y**2
      out(7) = dble(sngl(x))
      out(8) = dsign(1.0d+00,-0.0d0*x)
      k = 0
      do 10 i = 1, n
        if (i.eq.3) go to 10
        k = k + i
   10 continue
      out(9) = k
      out(10) = i
      return
      end
      function acc(v)
      real*8 acc, v, total
      save total
      total = total + v
      acc = total
      end
      subroutine t2(a, m, n, r)
      double precision a(m,n), r
      integer m, n
      r = 0.0d0
      do j = 1, n
        do i = m, 1, -2
          r = r + a(i,j)*i
          if (r .gt. 1.d3) exit
        enddo
      enddo
      call bump(r, 2)
      end
      subroutine bump(r, k)
      double precision r
      r = r + k
      end
"""


def test_translator_semantics():
    import f77toc
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "t.f")
        open(f, "w").write(SRC)
        tr = f77toc.Translator().run(f77toc.read_statements(f))
        c = os.path.join(d, "t.c")
        open(c, "w").write(tr.c_source(f))
        so = os.path.join(d, "t.so")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-shared", "-o", so, c, "-lm"])
        L = C.CDLL(so)
        x, y, n = C.c_double(1.7), C.c_double(0.3), C.c_int(5)
        out = np.zeros(10)
        L.t1_(C.byref(x), C.byref(y), C.byref(n), out.ctypes.data_as(C.c_void_p))
        assert out[0] == float(np.float32(0.01)) * 1.7 and out[0] != 0.01 * 1.7
        assert out[1] == 0.01 * 1.7
        assert out[2] == float(np.float32(0.1) * np.float32(3))
        assert out[3] == 3.0 and out[4] == 3.5
        assert out[5] == 1.7 * 1.7 - 0.3 * 0.3
        assert out[6] == float(np.float32(1.7))
        assert out[7] == -1.0                       # dsign sees the sign of a negative zero
        assert out[8] == 1 + 2 + 4 + 5 and out[9] == 6   # label 10 = next trip; the DO variable ends one step past
        L.acc_.restype = C.c_double
        v = C.c_double(2.5)
        assert L.acc_(C.byref(v)) == 2.5 and L.acc_(C.byref(v)) == 5.0      # save
        a = np.arange(1.0, 13.0)                    # a(3,4) column-major
        m, nn, r = C.c_int(3), C.c_int(4), C.c_double(0)
        L.t2_(a.ctypes.data_as(C.c_void_p), C.byref(m), C.byref(nn), C.byref(r))
        A = a.reshape(4, 3).T
        assert r.value == sum(A[i - 1, j] * i for j in range(4) for i in (3, 1)) + 2
