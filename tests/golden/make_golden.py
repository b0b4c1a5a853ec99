"""Extracts the reference's own golden vectors for the hot path into small fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Sources:
  helpers/helpers_test.go:167-176    SHA-512("x")
  snappy/hashes_test.go:57-104       golden hashes.yaml (+ SHA-512("") and SHA-512("bar\\n"))
  snappy/hashes_test.go:30-33        single fileHash document
  snappy/systemimage_test.go:104,117 SHA-512 of the 46-byte version string
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def main():
    hashes_test = (REF / "snappy/hashes_test.go").read_text()
    golden = hashes_test[hashes_test.index("Equals, `archive-sha512") + len("Equals, `"):]
    golden = golden[: golden.index("`)")]
    (OUT / "hashes_simple.yaml").write_text(golden)

    frag = hashes_test[hashes_test.index("var fileHashYaml = `") + len("var fileHashYaml = `"):]
    frag = frag[: frag.index("`")]
    (OUT / "filehash_fragment.yaml").write_text(frag)

    helpers_test = (REF / "helpers/helpers_test.go").read_text()
    x_digest = re.search(r'Equals, "([0-9a-f]{128})"', helpers_test).group(1)

    si_test = (REF / "snappy/systemimage_test.go").read_text()
    version_details = re.search(r"version_detail: (\S+)", si_test)
    si_digest = re.search(r'"([0-9a-f]{128})"', si_test).group(1)
    details = version_details.group(1) if version_details else "ubuntu=20141206,raw-device=20141206,version=77"

    empty = re.search(r"archive-sha512: ([0-9a-f]{128})", golden).group(1)
    bar = re.search(r"name: bin/bar\n  size: 4\n  sha512: ([0-9a-f]{128})", golden).group(1)
    kats = [
        {"source": "helpers/helpers_test.go:171-175", "message": "x", "sha512": x_digest},
        {"source": "snappy/hashes_test.go:89,101", "message": "", "sha512": empty},
        {"source": "snappy/hashes_test.go:95", "message": "bar\n", "sha512": bar},
        {"source": "snappy/systemimage_test.go:104,117", "message": details, "sha512": si_digest},
    ]
    (OUT / "sha512_kats.json").write_text(json.dumps(kats, indent=1) + "\n")
    print("wrote", [p.name for p in OUT.iterdir() if p.suffix in (".yaml", ".json")])


if __name__ == "__main__":
    main()
