#!/usr/bin/env python
"""Generate tests/golden/synth_hashes.json -- the drift pin of the "restatement-only" cases.

    python tests/golden/gen_synth_hashes.py

For every synthetic case of tests/synth_cases.py (exotic shapes, frame-length sweep, entropy edges, hostile
mutations -- everything FFmpeg cannot emit, where the CUDA path and the oracle are only ever compared with each other)
this records
    packets_sha256   the generated input packets (so a changed generator / encoder is told apart from a changed decoder)
    result_sha256    per packet: status word, byte count and PCM bytes, as the ORACLE decodes them today
    statuses         histogram of status codes
Both suites assert it: tests/test_oracle_golden.py (oracle == committed) and tests/test_gpu_parity.py (CUDA path ==
committed), so the oracle and the kernel cannot drift together unnoticed. Regenerate only when a case list changes, and
say why in the commit.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth_pin  # noqa: E402


def main():
    table = synth_pin.compute_with_oracle()
    path = os.path.join(HERE, 'synth_hashes.json')
    with open(path, 'w') as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print(len(table['cases']), 'cases,', table['packets'], 'packets ->', path)


if __name__ == '__main__':
    main()
