#!/usr/bin/env bash
# Single-box rehearsal of the multi-machine parameter server: two actor servers on loopback ports,
# then the driver placing its node actors round-robin over them (tcp://127.0.0.1:PORT).
set -euo pipefail
cd "$(dirname "$0")/../.."
ROUNDS=${ROUNDS:-5}
PORT_A=${PORT_A:-29101}
PORT_B=${PORT_B:-29102}
python examples/distributed/server.py --host 127.0.0.1 --port "$PORT_A" & a=$!
python examples/distributed/server.py --host 127.0.0.1 --port "$PORT_B" & b=$!
trap 'kill $a $b 2>/dev/null || true' EXIT
for port in "$PORT_A" "$PORT_B"; do          # wait until both servers accept connections
  for _ in $(seq 1 100); do
    python - "$port" <<'PY' && break || sleep 0.2
import socket, sys
s = socket.socket(); s.settimeout(0.2)
sys.exit(0 if s.connect_ex(("127.0.0.1", int(sys.argv[1]))) == 0 else 1)
PY
  done
done
python examples/distributed/mnist.py --remote-hosts "tcp://127.0.0.1:$PORT_A,tcp://127.0.0.1:$PORT_B" --rounds "$ROUNDS" --eval-interval 1
echo "distributed parameter-server rehearsal finished"
