"""Run the reference's UNMODIFIED main.py on the GPU backend.

    python tools/run_reference_main.py /path/to/qLDPC-branched-off [workdir]

``qldpc_b200.install_as_src()`` aliases this package as the reference's ``src`` package, then main.py is executed with
runpy from ``workdir`` (default: the reference checkout, whose codes/ and matrix_cache/ it then uses; a workdir without
codes/ gets them from ``qldpc_b200.codes.generate``).  matplotlib is optional: without it the plotting calls of
main.py:96-104 write nothing.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(reference_dir, workdir=None):
    import qldpc_b200
    qldpc_b200.install_as_src()
    workdir = workdir or reference_dir
    os.makedirs(workdir, exist_ok=True)
    if not os.path.isdir(os.path.join(workdir, "codes")):
        from qldpc_b200.codes.generate import generate_all
        generate_all(os.path.join(workdir, "codes"))
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        return runpy.run_path(os.path.join(reference_dir, "main.py"), run_name="__main__")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    run(os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else None)
