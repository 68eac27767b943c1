#!/usr/bin/env python
"""Run the REFERENCE'S OWN unit tests of the plugin (tests/compact/compress_fastpath_test.py,
compress_slowpath_test.py, prof_test.py -- unmodified, staged under the git-ignored baseline/_ref by
tools/stage_reference.sh in the build container) against THIS library: `xfuser.compact.*` and `xfuser.prof`
are aliased to compactfusion_b200 by the shim before the test modules are imported.  The drop-in claim in its
bluntest form: the tests the reference wrote for its Triton / eager-torch codecs, passing on our kernels.

    bash tools/stage_reference.sh            # build container (needs /root/reference)
    gpurun -- 'python tools/run_reference_tests.py > gpurun_out/ref_tests.log 2>&1'
    python tools/run_reference_tests.py --collect-only      # CPU: import / collection check only
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    tests = os.path.join(REF, "tests", "compact")
    if not os.path.isdir(tests):
        print(f"{tests} not staged: run tools/stage_reference.sh in the build container first")
        return 2
    sys.path.insert(0, ROOT)
    import compactfusion_b200.shim as shim
    # the staged copy also holds the reference's own xfuser/compact sources (for tools/ref_gpu_compare.py):
    # they must NOT shadow the shim, so only the tests directory's parent goes on sys.path, after the alias
    shim.install()
    assert sys.modules["xfuser.compact.fastpath"].__name__.startswith("compactfusion_b200")
    sys.path.insert(0, REF)  # `from tests.compact.compress_slowpath_test import ...` inside the fastpath test
    import pytest
    args = [tests, "-q", "-p", "no:cacheprovider", "--rootdir", REF, "-o", "python_files=*_test.py"] + sys.argv[1:]
    return pytest.main(args)


if __name__ == "__main__":
    sys.exit(main())
