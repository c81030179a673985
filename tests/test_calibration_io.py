"""CPU-only: calibration text IO in the reference's format (SaveCalibration / LoadCalibration, io.cc:570-700)."""
import numpy as np
import pytest

from badslam_b200 import calibration_io as IO


def test_files_have_the_reference_layout(tmp_path):
    base = str(tmp_path / "calib")
    d = np.array([525.25, 524.75, 319.5, 239.5], np.float32)
    c = np.array([540.0, 541.5, 320.75, 240.125], np.float32)
    cf = np.array([[0.001, -0.0025, 0.0], [1.5e-5, 0.125, 3.0]], np.float32)
    IO.write_calibration(base, d, c, 0.0312345678, cf)
    # "fx fy cx-0.5 cy-0.5" with ostream's default 6 significant digits, no trailing newline (io.cc:580-583)
    assert open(base + ".depth_intrinsics.txt").read() == "525.25 524.75 319 239"
    assert open(base + ".color_intrinsics.txt").read() == "540 541.5 320.25 239.625"
    # "w h" / a / one cfactor value per line, row by row, 8 significant digits (io.cc:609-619)
    lines = open(base + ".deformation.txt").read().split("\n")
    assert lines[0] == "3 2" and lines[1] == "0.031234568" and lines[-1] == ""
    # (-0.0025f is -0.00249999994...: at 8 digits iostream prints what is stored)
    assert lines[2:8] == ["0.001", "-0.0024999999", "0", "1.5e-05", "0.125", "3"]


def test_number_formatting_equals_iostream(tmp_path):
    """The writer must produce byte for byte what `ofstream << float` does (default precision for the intrinsics, precision(8)
    for the deformation file): checked against a C++ program that streams the same values."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no C++ compiler")
    rng = np.random.default_rng(3)
    d = (np.array([525, 525, 320, 240]) + rng.normal(0, 3, 4)).astype(np.float32)
    c = (np.array([1050, 1050, 640, 360]) + rng.normal(0, 3, 4)).astype(np.float32)
    cf = np.concatenate([5e-3 * rng.standard_normal(40), 10.0 ** rng.uniform(-9, 3, 20), [0.0, 1.0, -1e-5]]).astype(np.float32).reshape(7, 9)
    a = np.float32(0.0301234)
    base = str(tmp_path / "py")
    IO.write_calibration(base, d, c, a, cf)
    np.concatenate([d, c, [a], cf.reshape(-1)]).astype(np.float32).tofile(str(tmp_path / "values.bin"))
    src = tmp_path / "w.cpp"
    src.write_text(r'''
#include <fstream>
#include <vector>
int main(int argc, char** argv) {
  std::string dir = argv[1];
  std::ifstream in(dir + "/values.bin", std::ios::binary);
  std::vector<float> v(8 + 1 + 63);
  in.read(reinterpret_cast<char*>(v.data()), v.size() * sizeof(float));
  const char* names[2] = {"/cpp.depth_intrinsics.txt", "/cpp.color_intrinsics.txt"};
  for (int i = 0; i < 2; ++i) {   // io.cc:576-584
    std::ofstream f(dir + names[i], std::ios::out);
    const float* p = v.data() + 4 * i;
    f << p[0] << " " << p[1] << " " << (p[2] - 0.5) << " " << (p[3] - 0.5);
  }
  std::ofstream f(dir + "/cpp.deformation.txt", std::ios::out);   // io.cc:605-619
  f.precision(8);
  f << 9 << " " << 7 << std::endl;
  f << v[8] << std::endl;
  for (int i = 0; i < 63; ++i) f << v[9 + i] << std::endl;
  return 0;
}''')
    exe = tmp_path / "w"
    subprocess.check_call([gxx, "-std=c++17", str(src), "-o", str(exe)])
    subprocess.check_call([str(exe), str(tmp_path)])
    for suffix in (".depth_intrinsics.txt", ".color_intrinsics.txt", ".deformation.txt"):
        assert open(base + suffix).read() == open(str(tmp_path / "cpp") + suffix).read(), suffix


def test_round_trip_and_errors(tmp_path):
    rng = np.random.default_rng(0)
    base = str(tmp_path / "calib")
    d = (np.array([525, 525, 320, 240]) + rng.normal(0, 1, 4)).astype(np.float32)
    c = (np.array([540, 540, 320, 240]) + rng.normal(0, 1, 4)).astype(np.float32)
    cf = (5e-3 * rng.random((60, 80))).astype(np.float32)
    a = 0.0297
    IO.write_calibration(base, d, c, a, cf)
    d2, c2, a2, cf2 = IO.read_calibration(base, cf.shape)
    assert np.allclose(d2, d, rtol=1e-5) and np.allclose(c2, c, rtol=1e-5)       # 6 significant digits on disk
    assert abs(a2 - a) < 1e-8 and np.allclose(cf2, cf, rtol=1e-7, atol=0) and cf2.shape == cf.shape
    with pytest.raises(ValueError):       # the reference refuses a grid of another size (io.cc:676-680)
        IO.read_calibration(base, (30, 40))
    with pytest.raises(OSError):
        IO.read_calibration(str(tmp_path / "missing"))
    open(base + ".deformation.txt", "w").write("80 60\n0.03\n0.1\n")
    with pytest.raises(ValueError):
        IO.read_calibration(base, (60, 80))


class _FakeBA:
    """The five accessors SaveCalibration / LoadCalibration use, without a GPU."""

    def __init__(self):
        from badslam_b200.direct_ba import PinholeCamera4f
        self.P = PinholeCamera4f
        self.d, self.c = self.P(640, 480, [525, 525, 320, 240]), self.P(640, 480, [540, 540, 320, 240])
        self._a, self.cf = 0.01, np.full((4, 5), 0.002, np.float32)

    def depth_camera(self): return self.d
    def color_camera(self): return self.c
    def a(self): return self._a
    def cfactor_buffer(self, stream=None): return self.cf.copy()
    def SetDepthCamera(self, cam): self.d = cam
    def SetColorCamera(self, cam): self.c = cam
    def SetA(self, a): self._a = a
    def SetCFactorBuffer(self, cf, stream=None): self.cf = np.asarray(cf, np.float32).copy()


def test_save_and_load_through_the_directba_accessors(tmp_path):
    base = str(tmp_path / "calib")
    src, dst = _FakeBA(), _FakeBA()
    src.d = src.P(640, 480, [526.5, 524.25, 318.75, 241.0])
    src._a, src.cf = 0.03, (1e-3 * np.arange(20, dtype=np.float32)).reshape(4, 5)
    assert IO.SaveCalibration(src, base)
    assert IO.LoadCalibration(dst, base)
    assert np.allclose(dst.d.parameters, src.d.parameters) and np.allclose(dst.c.parameters, src.c.parameters)
    assert abs(dst._a - 0.03) < 1e-8 and np.allclose(dst.cf, src.cf, rtol=1e-7)
    assert (dst.d.width, dst.d.height) == (640, 480)
    assert not IO.LoadCalibration(dst, str(tmp_path / "missing"))       # false, like the reference, not an exception
