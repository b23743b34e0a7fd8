"""Record tests/golden/deterministic_signature.json: the payload of the REFERENCE's
scripts/deterministic_episode_signature.py, run against the reference build in oracle/_ref (whose sha256 is the
pin 0a8fe5cd...4d24 of SURVEY 8c).  The simulator package imports gymnasium / boto3 at import time; they are absent
in this image and never used by the script, so inert stand-ins are registered first.

    python tests/golden/make_golden_signature.py
"""

import hashlib
import json
import runpy
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def main():
    from oracle import reference

    assert reference.load() is not None, "build oracle/_ref first (make -f oracle/Makefile.ref)"
    for name in ["gymnasium", "gymnasium.spaces", "boto3", "botocore", "botocore.exceptions"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["gymnasium"].spaces = sys.modules["gymnasium.spaces"]
    for n in ["Box", "Discrete", "MultiDiscrete", "Space", "Dict"]:
        setattr(sys.modules["gymnasium.spaces"], n, type(n, (object,), {"__init__": lambda self, *a, **k: None}))
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore.exceptions"].NoCredentialsError = Exception
    g = runpy.run_path(str(reference.REF_ROOT / "scripts" / "deterministic_episode_signature.py"))
    payload = g["build_signature_payload"]()
    digest = hashlib.sha256(json.dumps(payload, sort_keys=True, separators=(",", ":")).encode()).hexdigest()
    assert digest == "0a8fe5cd26e34f712ba3035e386418fe428a5cdf268a92f7efd8658d5fef4d24", digest
    (HERE / "deterministic_signature.json").write_text(json.dumps(payload, sort_keys=True, indent=1))
    print(digest)


if __name__ == "__main__":
    main()
