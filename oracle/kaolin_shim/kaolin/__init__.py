"""Pure-torch stand-in for the five NVIDIA kaolin (pinned v0.13.0, reference Dockerfile:33) SPC ops
that PRBonn/SHINE_mapping's hot path calls.  TEST INFRASTRUCTURE ONLY: it exists so the *unmodified*
reference classes under /root/reference can be imported in the build container (kaolin is not installable
offline) to validate `oracle/shine_oracle.py` and to mint `tests/golden/*`.  Nothing in the product
package (`shine_mapping_b200/`) imports this.

kaolin's source is NOT vendored in the reference; semantics below follow kaolin's published docs and are
cross-pinned by how the reference consumes them (model/feature_octree.py:186-195,229 pins corner order).
"""
from . import ops  # noqa: F401
from . import render  # noqa: F401
