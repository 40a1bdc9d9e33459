"""The second Fortran 90 translator (oracle/f90toc_love.py: what the R/T pins lean on) on a small module with known answers,
checked against numpy's complex arithmetic: COMPLEX*16 under gcc's Fortran rules, whole arrays / sections / constructors /
MATMUL / RESHAPE scalarised with evaluate-then-store semantics, a pointer associated with a section, a rank-3 allocatable with
fixed leading extents, the derived type as a struct, an array-valued function, SELECT CASE, MERGE, a DO without control, an
internal procedure reading its host's local and its host's dummy argument."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GRT = """
module m_GRT
    use iso_c_binding
    implicit none
    integer, parameter :: dp=c_double
    integer, parameter :: nmode = 4
    complex*16,parameter::IC=(1d0,0d0)
    real(kind=dp),parameter::eps=1d-10
    type T_GRT
        integer nlayers
        real(kind=dp)  w
        real(kind=dp), dimension(:), allocatable :: vs
        real(kind=dp) :: root1(nmode)
    endtype T_GRT
contains
    complex*16 function csq(c,vel)
        implicit none
        real(kind=dp) c,vel
        csq=sqrt(dcmplx(1-(c/vel)**2))
    end function csq
end module m_GRT
"""

WORK = """
module work
    use m_GRT, only: dp, T_GRT, csq, IC
    implicit none
    complex*16,target,allocatable::R(:,:,:)
    complex*16::a22(2,2),b22(2,2),cs(0:1),la(2)
    complex*16,target::a44(4,4)
    complex*16,pointer::pp(:,:)
CONTAINS
    function twice(a)
        implicit none
        complex*16 twice(2,2)
        complex*16,intent(in)::a(2,2)
        twice=reshape([a(1,1),a(2,1),a(1,2),a(2,2)],[2,2])*2
    end function twice
    subroutine fill(c,GRT,k,out)
        implicit none
        real(kind=dp),intent(in)::c
        type(T_GRT), intent(in) :: GRT
        integer,intent(in)::k
        complex*16 out(12)
        integer i, j
        cs(0)=csq(c,GRT%vs(1)); cs(1)=csq(c,GRT%vs(GRT%nlayers))
        a22(:,1)=[IC,GRT%w*cs(1)]
        a22(:,2)=[-cs(0),a22(2,1)/cs(0)]
        b22=a22
        b22=b22/(2.*b22(2,1))          ! every element by the OLD b22(2,1)
        la=exp(-c*[cs(0),cs(1)])
        do i=1,2
            b22(:,i)=b22(:,i)*la(i)
        enddo
        allocate(R(2,2,0:GRT%nlayers))
        R = 0
        R(:,:,k)=matmul(a22,b22)
        a44=0.
        a44(1:2,3:4)=twice(R(:,:,k))
        pp=>a44(1:2,3:4)
        pp=-pp
        select case(k)
        case(1)
            a44(4,4)=merge(IC,-IC,c>1.)
        case(2)
            a44(4,4)=(0d0,3d0)
        end select
        j=0
        do
            j=j+1
            if(j>=3) exit
        enddo
        out(1:4)=[a22(1,1),a22(2,1),a22(1,2),a22(2,2)]
        out(5:8)=[b22(1,1),b22(2,1),b22(1,2),b22(2,2)]
        out(9:12)=[a44(1,3),a44(2,4),a44(4,4),dcmplx(j+aimag(la(2)))]
        deallocate(R)
    end subroutine fill
end module work
subroutine host(x,res)
    implicit none
    real*8 x,res,acc
    acc=1d0
    call bump
    call bump
    res=acc
contains
    subroutine bump
        acc=acc*x+1d0
    end subroutine bump
end subroutine host
"""


def test_translator_semantics():
    import f90toc_love as T
    with tempfile.TemporaryDirectory() as d:
        g, w = os.path.join(d, "GRT.f90"), os.path.join(d, "work.f90")
        open(g, "w").write(GRT)
        open(w, "w").write(WORK)
        c = os.path.join(d, "t.c")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "f90toc_love.py"), g, "host=acc:real,x:real*@host", w, c])
        src = open(c).read().replace("static void fill_", "void fill_").replace("static void host_", "void host_")
        open(c, "w").write(src)
        so = os.path.join(d, "t.so")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-fcx-fortran-rules", "-ffp-contract=off", "-shared", "-o", so, c, "-lm"])
        L = C.CDLL(so)

        class TGRT(C.Structure):
            _fields_ = [("nlayers", C.c_int), ("w", C.c_double), ("vs", C.POINTER(C.c_double)), ("vs_d1", C.c_int), ("vs_l1", C.c_int),
                        ("root1", C.c_double * 4)]
        vs = np.array([3.0, 3.5, 2.2])
        grt = TGRT(3, 1.7, vs.ctypes.data_as(C.POINTER(C.c_double)), 3, 1)
        for k, cval in ((1, 2.6), (2, 0.9)):
            out = np.zeros(12, complex)
            cc, kk = C.c_double(cval), C.c_int(k)
            L.fill_(C.byref(cc), C.byref(grt), C.byref(kk), out.ctypes.data_as(C.c_void_p))
            cs0, cs1 = np.sqrt(complex(1 - (cval / 3.0) ** 2)), np.sqrt(complex(1 - (cval / 2.2) ** 2))
            a = np.array([[1, -cs0], [1.7 * cs1, 1.7 * cs1 / cs0]])
            la = np.exp(-cval * np.array([cs0, cs1]))
            b = a / (2.0 * a[1, 0]) * la[None, :]
            r = a @ b
            assert np.allclose(out[:4], a.T.ravel(), rtol=1e-14, atol=0) and np.allclose(out[4:8], b.T.ravel(), rtol=1e-13, atol=1e-300)
            assert np.isclose(out[8], -2 * r[0, 0], rtol=1e-13) and np.isclose(out[9], -2 * r[1, 1], rtol=1e-13)     # twice(), then pp = -pp
            assert out[10] == ((1 if cval > 1 else -1) if k == 1 else 3j)                                             # SELECT CASE, MERGE
            assert np.isclose(out[11], 3 + la[1].imag, rtol=1e-13)                                                    # DO ... EXIT: j = 3
        x, res = C.c_double(1.5), C.c_double(0)
        L.host_(C.byref(x), C.byref(res))
        assert res.value == (1.0 * 1.5 + 1) * 1.5 + 1                                                                  # the internal procedure saw acc and x
