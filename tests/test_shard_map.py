"""CPU tier: the tile / shard index arithmetic every covariance kernel uses (conan_slam_b200/csrc/shard_map.h,
plain C++) checked exhaustively on the host — all owned tiles in row-major order for world sizes 1-8, tile
counts up to a 500 000-row covariance."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_tile_enumerates_the_owned_triangle(tmp_path):
    exe = str(tmp_path / "shard_map_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "shard_map_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shard map ok" in out.stdout
