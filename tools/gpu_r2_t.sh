#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -rf /tmp/dbg_ckpt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tools/dbg_dist_resume.py > gpurun_out/t_dbg.log 2>&1
echo "rc=$?" >> gpurun_out/t_dbg.log
grep -v "^\[W\|OMP_NUM\|\*\*\*\*" gpurun_out/t_dbg.log | tail -60
