"""The reference's OWN test file (/root/reference/tests/test_array.py), run UNCHANGED against the product
package: `import arraylib as al` inside it resolves to the shim at the repo root, i.e. to the CUDA library.

The file is not copied into this repo. `make -C oracle reftests` (part of `__graft_entry__.build()`)
byte-compiles it from where it lies into oracle/_ref/test_array.marshal -- the same arrangement as the compiled
reference library next to it -- and this module executes that code object in its own namespace, so pytest
collects the reference's test functions with their own parametrisation (822 cases, among them the 2 x 360
per-index get/set cases of test_array.py:255-270). tests/test_gpu_reference_suite.py remains as the readable
restatement of the same cases."""
import marshal
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

_ARTEFACT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "test_array.marshal")


def _load():
    if not os.path.exists(_ARTEFACT):
        return None, "oracle/_ref/test_array.marshal is missing (run `make -C oracle reftests` where /root/reference exists)"
    with open(_ARTEFACT, "rb") as f:
        header, payload = f.read().split(b"\n", 1)
    version = header.split()[0].decode()
    if version != f"{sys.version_info[0]}.{sys.version_info[1]}":
        return None, f"compiled for python {version}, running {sys.version_info[0]}.{sys.version_info[1]}"
    return marshal.loads(payload), None


_code, _why = _load()
if _code is not None:
    exec(_code, globals())  # defines the reference's test_* functions here, decorators and all
else:
    def test_reference_file_is_available():
        pytest.fail(_why)


@pytest.fixture(autouse=True, scope="module")
def _bound_to_the_gpu(al):  # fails (does not skip) when there is no device
    return al
