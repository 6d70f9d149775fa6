#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 80 python -m pytest tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -2
