"""Every BASELINE config at FULL size against the oracle's committed digests (tests/golden/families_c{1..5}.json).

The digests were made offline on the CPU by tests/golden/make_families_golden.py: the reference's own libdivsufsort64
(oracle/_ref) for the suffix array + the oracle's restatement of SearchDuplications::run and the post-steps
(src/bin/asgart.rs:137-258, :33-96) + its JSON exporter. Here the CUDA path must reproduce, through the C ABI,
  * prepare_data's chunk list and fragment map            (chunks_sha256, map)
  * divsufsort64's suffix array                            (sa_fingerprint, computed on the device)
  * the families, order included, and the counters         (families_sha256, probes / searched / matches)
  * the JSON text JSONExporter::save would write           (json_sha256)
C4 is the configuration the bench line is quoted on (3.1 Gbp, ~115 GB of HBM, ~10 s of GPU time)."""
import hashlib
import json
import os

import numpy as np
import pytest

import asgart_b200 as ab
from asgart_b200.api import families_digest, sa_fingerprint_host
from bench import CONFIG_FLAGS, make_workload

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(config):
    with open(os.path.join(GOLD, f"families_c{config}.json")) as f:
        return json.load(f)


def test_fingerprint_device_equals_host():
    rng = np.random.default_rng(3)
    t = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=300_001)
    with ab.Context(0) as ctx:
        ctx.load_strand(np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)]))
        for bits in (32, 64):
            ctx.set_index_bits(bits)
            ctx.build_index()
            assert ctx.sa_fingerprint() == sa_fingerprint_host(ctx.download_sa())


@pytest.mark.parametrize("config", [1, 2, 3, 5, 4])
def test_full_size_config_equals_oracle_digest(config):
    gold = _gold(config)
    assert gold["full_size"] and gold["flags"] == CONFIG_FLAGS[config]
    st, prep = make_workload(config, 0)
    assert prep.n1 - 1 == gold["strand_bp"]
    assert [list(m) for m in prep.map] == gold["map"]
    ch = np.array(prep.chunks, dtype="<u8").reshape(-1, 2)
    assert len(prep.chunks) == gold["n_chunks"] and hashlib.sha256(ch.tobytes()).hexdigest() == gold["chunks_sha256"]
    with ab.Context(0) as ctx:
        ctx.load_strand(prep.strand)
        ctx.build_index()
        assert ctx.sa_fingerprint() == gold["sa_fingerprint"], "suffix array differs from the reference's divsufsort64"
        ctx.reset_stats()
        fam = ctx.search(prep.chunks, st, ab.POST_ALL)
        s = ctx.stats()
    assert (fam.n_families, len(fam.sds)) == (gold["families"], gold["duplicons"])
    assert fam.digest() == gold["families_sha256"]
    for ours, theirs in (("n_probes", "probes"), ("n_searched", "searched"), ("n_skipped_n", "skipped_n"),
                         ("n_skipped_card", "skipped_card"), ("n_matches", "matches")):
        assert s[ours] == gold["counters"][theirs], ours
    js = prep.to_json(st, fam)
    assert len(js) == gold["json_bytes"] and hashlib.sha256(js.encode()).hexdigest() == gold["json_sha256"]
