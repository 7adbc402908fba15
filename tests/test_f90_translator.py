"""Unit tests of the Fortran-90-subset -> C++ translator that pins the oracle (oracle/f90_to_cpp.py).

The translator is trusted to carry the reference's semantics, so its own rules are tested on small
hand-checkable programs: operator precedence and association, unary minus, integer division and
real->integer truncation, nint, implicit typing, mixed-mode arithmetic, `**` expansion, parameter
initialisers evaluated in fp32, `use ... only` scoping, by-reference arguments, whole-array /
section assignments in column-major order, where / elsewhere, forall, do loops, sum() order,
direct-access writes and print capture."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("f90_to_cpp", os.path.join(ROOT, "oracle", "f90_to_cpp.py"))
f90 = importlib.util.module_from_spec(spec)
spec.loader.exec_module(f90)

SRC = """
module mo_a
  integer, parameter :: nx = 4, ny = 3
  integer, parameter :: half = 0.5*7          ! real -> integer truncation: 3
  real, parameter    :: third = 1./3
  real :: c1 = 273.15-1.7                      ! folded in single precision
  real :: c2 = -0.1/24./3600.
  real, dimension(nx,ny) :: a, b, m
  real, dimension(ny,5)  :: s
  real, dimension(4) :: pe = (/1.5, 2.5, 3.5, 4.5/)
  integer :: jj = 0
end module mo_a

module mo_b
  use mo_a, only: nx, ny
  real, dimension(nx,ny) :: acc
  real :: hidden = 7.0
end module mo_b

subroutine prec(x, y, z, r)
  implicit none
  real :: x, y, z
  real, dimension(8) :: r
  r(1) = -x*y**2          ! -(x*(y*y))
  r(2) = x - y - z        ! (x-y)-z
  r(3) = x/y/z            ! (x/y)/z
  r(4) = -x + y           ! (-x)+y
  r(5) = x*y/z*x          ! ((x*y)/z)*x
  r(6) = (x+y)**4         ! t=(x+y)*(x+y); t*t
  r(7) = 2.**2 + 3.0**2
  r(8) = x + y*z - x/y
end subroutine

subroutine ints(i, k, x, out)
  ! implicit typing: i,k integer, x real, out declared
  real, dimension(8) :: out
  out(1) = i/k            ! integer division truncates toward zero
  out(2) = (0-i)/k
  out(3) = nint(x)        ! half away from zero
  out(4) = nint(-x)
  out(5) = int(x*3)
  out(6) = mod(i, k)
  out(7) = i/float(k)
  n = x*4                 ! implicit integer n: truncation on assignment
  out(8) = n
end subroutine

subroutine arrays(t)
  use mo_a
  use mo_b, only: acc
  real, dimension(nx,ny) :: t
  a = t*2. + 1.
  b(:,2) = a(:,1) - t(:,3)
  forall (i=1:nx)
     s(:,2) = t(i,:)*pe(2)      ! last i wins (statement executes for i = 1..nx in order)
  end forall
  where (t > 5.) m = 1.
  where (t <= 5.) m = -1.
  where (a(:,:) >= 10.0)
     acc = a
     b = 0.0
  elsewhere
     acc = -a
  end where
  do j = 1, ny
     if (j == 2) then
        acc(1,j) = 100.
     else if (j == 3) then
        acc(1,j) = 200.
     else
        acc(1,j) = 300.
     end if
  end do
  jj = 0
  do i = 1, nx
     jj = jj + i
  end do
end subroutine

subroutine callee(v, n, w)
  real :: v
  real, dimension(3) :: w
  v = v + n
  n = n + 1
  w(2) = v
end subroutine

subroutine caller(out)
  use mo_b, only: hidden
  real, dimension(6) :: out
  real, dimension(3) :: w
  x = 1.5; k = 2
  call callee(x, k, w)
  out(1) = x; out(2) = k; out(3) = w(2)
  call callee(x, 10, w)           ! literal actual argument: a temporary
  out(4) = x
  out(5) = hidden                  ! imported
  c1 = 5.0                         ! NOT imported here: an implicit local, the module variable keeps its value
  out(6) = c1
end subroutine

subroutine sums(t, r)
  use mo_a, only: nx, ny
  real, dimension(nx,ny) :: t
  real, dimension(3) :: r
  integer, dimension(5) :: cnt = (/1,2,3,4,5/)
  r(1) = sum(t)
  r(2) = sum(cnt(1:3))
  r(3) = sum(t)/(nx*ny) - 273.15
  write(22, rec=2) t/2
  print *, 'text', r(1), nx
end subroutine
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("f90")
    src = d / "t.f90"
    src.write_text(SRC)
    cpp = d / "t.cpp"
    cpp.write_text(f90.translate(str(src)))
    so = d / "libt.so"
    subprocess.run(["g++", "-O3", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-w", "-o", str(so), str(cpp)],
                   check=True)
    return C.CDLL(str(so))


def f32(x):
    return np.float32(x)


def call(lib, name, *args):
    conv = []
    for a in args:
        if isinstance(a, np.ndarray):
            conv.append(a.ctypes.data_as(C.c_void_p))
        else:
            conv.append(C.byref(a))
    getattr(lib, "f_" + name)(*conv)


def test_precedence_and_association(lib):
    x, y, z = f32(1.7), f32(-2.3), f32(0.37)
    r = np.zeros(8, np.float32)
    call(lib, "prec", C.c_float(x), C.c_float(y), C.c_float(z), r)
    t = f32(x + y) * f32(x + y)
    want = [-(x * f32(y * y)), f32(x - y) - z, f32(x / y) / z, f32(-x) + y, f32(f32(x * y) / z) * x, t * t,
            f32(13.0), f32(f32(x + f32(y * z)) - f32(x / y))]
    assert np.array_equal(r, np.array(want, np.float32))


def test_integer_semantics_and_implicit_typing(lib):
    r = np.zeros(8, np.float32)
    call(lib, "ints", C.c_int(7), C.c_int(2), C.c_float(2.5), r)
    assert r.tolist() == [3.0, -3.0, 3.0, -3.0, 7.0, 1.0, 3.5, 10.0]


def test_module_initialisers_are_single_precision(lib):
    assert C.c_int.in_dll(lib, "f_half").value == 3
    assert f32(C.c_float.in_dll(lib, "f_third").value) == f32(1.0) / f32(3)
    assert f32(C.c_float.in_dll(lib, "f_c1").value) == f32(273.15) - f32(1.7)
    assert f32(C.c_float.in_dll(lib, "f_c2").value) == f32(f32(-0.1) / f32(24.0)) / f32(3600.0)


def test_array_statements(lib):
    nx, ny = 4, 3
    t = (np.arange(nx * ny, dtype=np.float32).reshape(ny, nx) * f32(1.1)).astype(np.float32)   # C [j][i] == Fortran (i,j)
    call(lib, "arrays", t)
    get = lambda n, shape: np.ctypeslib.as_array((C.c_float * int(np.prod(shape))).in_dll(lib, "f_" + n)).reshape(shape)
    a = t * f32(2) + f32(1)
    assert np.array_equal(get("a", (ny, nx)), a)
    assert np.array_equal(get("m", (ny, nx)), np.where(t > 5, 1, -1).astype(np.float32))
    b = get("b", (ny, nx))
    assert np.array_equal(b[1][a[1] < 10], (a[0] - t[2])[a[1] < 10])          # b(:,2) kept where a < 10
    assert np.all(b[a >= 10] == 0)
    s = get("s", (5, ny))
    assert np.array_equal(s[1], t[:, nx - 1] * f32(2.5))
    acc = get("acc", (ny, nx)).copy()
    assert acc[0, 0] == 300 and acc[1, 0] == 100 and acc[2, 0] == 200
    acc[:, 0] = np.where(a >= 10, a, -a)[:, 0]
    assert np.array_equal(acc, np.where(a >= 10, a, -a))
    assert C.c_int.in_dll(lib, "f_jj").value == 10


def test_by_reference_arguments_and_use_only_scoping(lib):
    out = np.zeros(6, np.float32)
    C.c_float.in_dll(lib, "f_c1").value = 1.25
    call(lib, "caller", out)
    assert out.tolist() == [3.5, 3.0, 3.5, 13.5, 7.0, 5.0]
    assert C.c_float.in_dll(lib, "f_c1").value == 1.25        # untouched: c1 was an implicit local in `caller`


def test_sum_order_write_and_print(lib):
    nx, ny = 4, 3
    rng = np.random.default_rng(0)
    t = rng.uniform(250, 300, (ny, nx)).astype(np.float32)
    r = np.zeros(3, np.float32)
    lib.f90_out_reset()
    call(lib, "sums", t, r)
    acc = f32(0)
    for v in t.ravel():                                        # array element order, fp32 accumulator
        acc = f32(acc + v)
    assert r[0] == acc and r[1] == 6.0
    assert r[2] == f32(f32(acc / f32(12)) - f32(273.15))
    lib.f90_out_nrecs.restype = C.c_size_t
    assert lib.f90_out_nrecs() == 1
    u, rec, n = C.c_int(), C.c_int(), C.c_size_t()
    lib.f90_out_rec_info(C.c_size_t(0), C.byref(u), C.byref(rec), C.byref(n))
    assert (u.value, rec.value, n.value) == (22, 2, 12)
    lib.f90_out_rec_data.restype = C.POINTER(C.c_float)
    data = np.ctypeslib.as_array(lib.f90_out_rec_data(C.c_size_t(0)), shape=(12,))
    assert np.array_equal(data, (t / f32(2)).ravel())
    lib.f90_print_nvals.restype = C.c_size_t
    lib.f90_print_data.restype = C.POINTER(C.c_double)
    assert lib.f90_print_nvals() == 2
    vals = np.ctypeslib.as_array(lib.f90_print_data(), shape=(2,))
    assert vals[0] == float(acc) and vals[1] == 4.0


def test_unsupported_constructs_raise():
    for bad in ("subroutine s(x)\n real :: x\n x = x**0.5\nend subroutine\n",
                "subroutine s(x)\n implicit none\n real :: x\n y = x\nend subroutine\n",
                "subroutine s(x)\n real :: x\n goto 10\nend subroutine\n"):
        p = "/tmp/_bad.f90"
        open(p, "w").write(bad)
        with pytest.raises(f90.F90Error):
            f90.translate(p)
