set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -12
timeout 300 python tools/h2d_probe.py 2>&1 | tail -12
