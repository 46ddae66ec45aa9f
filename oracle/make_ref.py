"""Recipe for oracle/_ref/: pack the reference's own Python files, byte for byte, into one archive.

The reference is pure Python (no C/C++ to compile), so the "build" of its hot path is an archive of the files
themselves: `oracle/_ref/reference_py.tar.gz`.  The directory is git-ignored (nothing of the reference enters the
history) but not gpurun-ignored, so the archive travels to the GPU box like the built `.so` files do.  There it gives
  * `bench.py --impl reference` / `cpu_baseline` the UNMODIFIED reference `ICP()` / `update_scan()` to time
    (`kind: "reference"`), instead of the oracle port;
  * `tests/test_gpu_slam_unmodified.py` the unmodified `slam.py` to drive against the drop-in shim.
Run in the build container only (`python oracle/make_ref.py`; `__graft_entry__.build()` calls it when
/root/reference exists).  TEST INFRASTRUCTURE ONLY -- the product never reads it.
"""
import io
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "reference_py.tar.gz")
FILES = ["slam.py", "config.yaml", "teapot.csv", "requirements.txt",
         "utilities/__init__.py", "utilities/icp.py", "utilities/mapping.py", "utilities/features.py",
         "utilities/pose_graph.py", "services/__init__.py", "services/lidar_service.py", "services/imu_service.py",
         "demos/teapot_icp_demo.py"]


def main():
    if not os.path.isdir(REF):
        print(f"{REF} is absent: nothing to pack (the GPU box uses the archive made in the build container)")
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz", compresslevel=6) as tar:
        for rel in FILES:
            path = os.path.join(REF, rel)
            if not os.path.exists(path):
                continue
            info = tar.gettarinfo(path, arcname=rel)
            info.mtime = 0                                  # reproducible archive
            info.uid = info.gid = 0
            info.uname = info.gname = ""
            with open(path, "rb") as f:
                tar.addfile(info, f)
    data = buf.getvalue()
    if not (os.path.exists(OUT) and open(OUT, "rb").read() == data):
        with open(OUT, "wb") as f:
            f.write(data)
    print(f"wrote {OUT} ({len(data)} bytes, {len(FILES)} files)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
