"""The Fortran 90 -> C translator (oracle/f90toc.py) on a small module with known answers: what it adds to f77toc's subset and
the fm2d pin leans on -- module variables, kind parameters and attribute declarations, allocatable arrays with lower bounds
(ALLOCATE / DEALLOCATE / ALLOCATED), array = array with reallocation of the left-hand side, the derived type and its component
references, DO WHILE / EXIT / CYCLE, NINT / FLOOR, continuation lines, parameters local to a unit, a synthetic subroutine made
of statement ranges."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GLOBALP = """
MODULE globalp
use iso_c_binding
IMPLICIT NONE
INTEGER, PARAMETER :: i10= c_double
INTEGER, PARAMETER :: i5=SELECTED_REAL_KIND(5,10)
INTEGER, SAVE :: nn, mm
REAL(KIND=i10), SAVE :: scale
REAL(KIND=i10), DIMENSION (:,:), ALLOCATABLE :: grid, copy   ! set by the units below
INTEGER, DIMENSION (:), ALLOCATABLE :: tags
REAL(KIND=i10), PARAMETER :: third=0.3333333
END MODULE globalp
"""

WORK = """
MODULE work
USE globalp
IMPLICIT NONE
INTEGER cnt
TYPE backpointer
   INTEGER :: px,pz
END TYPE backpointer
TYPE(backpointer), DIMENSION (:), ALLOCATABLE :: btg
!$omp threadprivate (cnt,btg)
CONTAINS

SUBROUTINE fill(a,b,out)
IMPLICIT NONE
INTEGER :: i,j,a,b
REAL(KIND=i10), DIMENSION (8) :: out
REAL(KIND=i5), PARAMETER :: eps=1.0e-3
TYPE(backpointer) :: t
IF(ALLOCATED(grid)) DEALLOCATE(grid)
ALLOCATE(grid(0:a,-1:b), STAT=cnt)
DO i=0,a
   DO j=-1,b
      grid(i,j)=10*i+j + &      ! a continuation line
                0.5
   ENDDO
ENDDO
copy=grid                       ! the allocatable left-hand side takes the shape AND the lower bounds
out(1)=copy(a,b)
out(2)=copy(0,-1)
out(3)=third*3
out(4)=NINT(2.5)+FLOOR(-0.5)+eps
IF(.NOT.ALLOCATED(btg))ALLOCATE(btg(3))
btg(1)%px=4
btg(1)%pz=7
btg(2)=btg(1)
t=btg(2)
btg(3)%px=t%pz+btg(2)%px
out(5)=btg(3)%px
nn=0
mm=0
DO WHILE(nn.lt.100)
   nn=nn+1
   IF(nn/2*2.EQ.nn)CYCLE
   IF(nn.GT.9)EXIT
   mm=mm+nn
ENDDO
out(6)=mm
out(7)=nn
scale=2.0*out(1)/REAL(3)
out(8)=scale
CALL bump
END SUBROUTINE fill

SUBROUTINE bump
IMPLICIT NONE
cnt=cnt+5
grid=-1
END SUBROUTINE bump

SUBROUTINE host(n,res)
IMPLICIT NONE
INTEGER :: n,k,acc
REAL(KIND=i10) :: res
LOGICAL :: unused
acc=0
k=1
acc=acc+1000
IF(n.GT.0)THEN
   DO k=1,n
      acc=acc+k
   ENDDO
ENDIF
acc=acc*2
res=acc
END SUBROUTINE host
END MODULE work
"""


def test_translator_semantics():
    import f90toc
    with tempfile.TemporaryDirectory() as d:
        g, w = os.path.join(d, "globalp.f90"), os.path.join(d, "work.f90")
        open(g, "w").write(GLOBALP)
        open(w, "w").write(WORK)
        mod = f90toc.Module()
        mod.read(f90toc.read_free_form(g))
        tr = f90toc.Translator90(mod)
        stmts = f90toc.read_free_form(w)
        tr.run(stmts)
        # a synthetic unit from two statement ranges of `host`: its declarations, without `acc=acc+1000` and `acc=acc*2`
        tr.run_ranges(stmts, host="host", name="part", args=["n", "res"], extra_decl=[],
                      ranges=[("acc=0", "k=1"), ("if(n.gt.0)then", "endif"), ("res=acc", "res=acc")])
        c = os.path.join(d, "t.c")
        src = tr.c_source([g, w]).replace("static void", "void")     # (the units are file-local for the fm2d driver)
        src += "\nint get_cnt(void) { return cnt; }\ndouble get_grid(int k) { return grid[k]; }\nint get_shape(int k) { int v[4] = {copy_d1, copy_d2, copy_l1, copy_l2}; return v[k]; }\n"
        open(c, "w").write(src)
        so = os.path.join(d, "t.so")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-shared", "-o", so, c, "-lm"])
        L = C.CDLL(so)
        a, b = C.c_int(3), C.c_int(2)
        out = np.zeros(8)
        L.fill_(C.byref(a), C.byref(b), out.ctypes.data_as(C.c_void_p))
        assert out[0] == 10 * 3 + 2 + 0.5 and out[1] == 10 * 0 - 1 + 0.5          # lower bounds 0 and -1, continuation line
        assert [L.get_shape(k) for k in range(4)] == [4, 4, 0, -1]                 # copy = grid: shape and lower bounds taken over
        assert out[2] == float(np.float32(0.3333333)) * 3                          # a default-real literal stays single
        assert out[3] == float(np.float32(3 - 1) + np.float32(1.0e-3))             # NINT(2.5) = 3, FLOOR(-0.5) = -1; integer + real*4 is real*4
        assert out[4] == 7 + 4                                                      # structure assignment and components
        assert out[5] == 1 + 3 + 5 + 7 + 9 and out[6] == 11                        # CYCLE on even, EXIT at 11
        assert out[7] == 2.0 * out[0] / 3.0
        L.get_grid.restype = C.c_double
        assert L.get_cnt() == 5 and all(L.get_grid(k) == -1.0 for k in range(16))  # STAT=cnt zeroed it; whole-array scalar assignment
        n, res = C.c_int(4), C.c_double(0)
        L.host_(C.byref(n), C.byref(res))
        assert res.value == (1000 + 10) * 2
        L.part_(C.byref(n), C.byref(res))
        assert res.value == 10.0
