"""Install the UNMODIFIED reference into ``baseline/_ref`` (git-ignored, travels with gpurun).

``pip install /root/reference`` fails as shipped — the tree has neither setup.py nor pyproject.toml (recorded in
DESIGN.md).  Following the task's recipe for that case, the tree is copied to /tmp, a packaging-only ``setup.py`` is
generated next to the (untouched) sources, and pip installs it with ``--no-deps --no-index --no-build-isolation
--target baseline/_ref``.  Only what the headline experiment needs is packaged: ``fedml_api``, ``fedml_core``,
``fedml_experiments`` and the SEA concept pools / change-point matrices under ``data/``.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FDB_REFERENCE_SRC", "/root/reference")
TMP = "/tmp/fdb_ref_src"
DST = os.path.join(HERE, "_ref")

SETUP = '''
import os
from setuptools import setup

def pkgs(root):
    out = []
    for d, _, files in os.walk(root):
        if "__pycache__" in d or any("." in part for part in d.split(os.sep)):
            continue
        if any(f.endswith((".py", ".sh")) for f in files):
            out.append(d.replace(os.sep, "."))
    return out

def data_files(root, exts):
    out = {}
    for d, _, files in os.walk(root):
        keep = [f for f in files if f.endswith(exts)]
        if keep and not any("." in part for part in d.split(os.sep)):
            out.setdefault(d.replace(os.sep, "."), []).extend(keep)
    return out

packages = pkgs("fedml_api") + pkgs("fedml_core") + pkgs("fedml_experiments") + ["data", "data.sea", "data.changepoints"]
package_data = data_files("fedml_experiments", (".sh", ".py"))
package_data["data.sea"] = ["concept1.csv", "concept2.csv", "concept3.csv", "concept4.csv"]
package_data["data.changepoints"] = [f for f in os.listdir("data/changepoints") if f.endswith(".cp")]
setup(name="feddrift-reference", version="0.0.0", packages=sorted(set(packages)), package_data=package_data,
      include_package_data=True, zip_safe=False)
'''


def main() -> int:
    if os.path.isdir(os.path.join(DST, "fedml_api")):
        print("reference already installed at", DST)
        return 0
    if not os.path.isdir(SRC):
        print("reference source not found:", SRC)
        return 1
    shutil.rmtree(TMP, ignore_errors=True)
    os.makedirs(TMP)
    for sub in ("fedml_api", "fedml_core", "fedml_experiments"):
        shutil.copytree(os.path.join(SRC, sub), os.path.join(TMP, sub),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "pretrained"))
    os.makedirs(os.path.join(TMP, "data"))
    shutil.copytree(os.path.join(SRC, "data", "changepoints"), os.path.join(TMP, "data", "changepoints"))
    os.makedirs(os.path.join(TMP, "data", "sea"))
    for f in os.listdir(os.path.join(SRC, "data", "sea")):
        if f.startswith("concept") and f.endswith(".csv"):
            shutil.copy(os.path.join(SRC, "data", "sea", f), os.path.join(TMP, "data", "sea", f))
    with open(os.path.join(TMP, "setup.py"), "w") as fh:
        fh.write(SETUP)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DST, TMP]
    print(" ".join(cmd))
    rc = subprocess.call(cmd)
    return rc


if __name__ == "__main__":
    sys.exit(main())
