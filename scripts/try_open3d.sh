#!/usr/bin/env bash
# Reproducible attempt to obtain the wheel the reference's TSDF half IS (open3d==0.17.0, /root/reference/requirements.txt:15).
# Run in the dev container and on the gpurun box; the transcript is committed under profiles/ either way.
# If the wheel ever becomes importable, tests/golden/make_tsdf_golden.py turns it into committed fixtures.
out="${1:-gpurun_out/open3d_attempt.log}"
mkdir -p "$(dirname "$out")"
{
  echo "== $(date -u +%FT%TZ) host=$(hostname) =="
  echo "\$ python -c 'import open3d'"
  python -c 'import open3d; print("open3d", open3d.__version__)' 2>&1 | tail -2
  echo "\$ ls /opt/wheelhouse | grep -i open3d"
  ls /opt/wheelhouse 2>/dev/null | grep -i open3d || echo "(none)"
  echo "\$ find / -iname 'open3d*' (site-packages, wheels)"
  find / -xdev \( -iname 'open3d*' -o -iname 'Open3D*' \) -not -path '*/proc/*' 2>/dev/null | head -5 || true
  echo "\$ pip download open3d==0.17.0 --no-deps -d /tmp/o3d (index, 20 s timeout)"
  timeout 60 python -m pip download open3d==0.17.0 --no-deps -d /tmp/o3d --timeout 5 --retries 0 2>&1 | tail -4
  echo "\$ pip install --no-index --find-links /opt/wheelhouse open3d==0.17.0"
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/o3d_target open3d==0.17.0 2>&1 | tail -3
  echo "\$ pip install open3d==0.17.0 (index)"
  timeout 60 python -m pip install --target /tmp/o3d_target open3d==0.17.0 --timeout 5 --retries 0 2>&1 | tail -3
  echo "\$ curl -sI https://pypi.org/simple/open3d/ (network probe)"
  timeout 10 curl -sI --max-time 8 https://pypi.org/simple/open3d/ 2>&1 | head -1 || echo "no route"
  echo "== verdict: $(PYTHONPATH=/tmp/o3d_target python -c 'import open3d; print("open3d importable", open3d.__version__)' 2>/dev/null || echo 'open3d NOT obtainable here') =="
} > "$out" 2>&1
cat "$out"
