#!/usr/bin/env bash
# Single-box smoke test: one server + 4 workers on localhost.
set -e
cd "$(dirname "$0")/../../.."
export BYZPY_HMAC_SECRET=${BYZPY_HMAC_SECRET:-local-test-secret}
python examples/ps/remote_tcp/ps_node.py server &
srv=$!
pids=()
for w in w0 w1 w2 w3; do
  python examples/ps/remote_tcp/ps_node.py worker --id $w &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
wait $srv
echo "tcp parameter-server smoke test finished"
