"""How the fixtures in this directory were made (run in the build container, where /root/reference exists).

RLdata500.csv.gz / RLdata10000.csv.gz: the reference's example DATA files (examples/*.csv, the RLdata sets of the R
package RecordLinkage with ground-truth ent_id), gzip -9, byte for byte.  They are inputs of BASELINE.json
configs[0] and configs[1]; the GPU box has no /root/reference, so the tests read these copies.
"""
import gzip
import shutil

for name in ("RLdata500", "RLdata10000"):
    with open(f"/root/reference/examples/{name}.csv", "rb") as f, \
            gzip.open(f"tests/golden/{name}.csv.gz", "wb", compresslevel=9) as g:
        shutil.copyfileobj(f, g)
