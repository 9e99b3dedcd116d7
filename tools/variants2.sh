cd $GRAFT_REPO_ROOT
timeout 100 python tools/mini_train.py 2>&1 | grep -E "run|Error" | head -4
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
