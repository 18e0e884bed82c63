"""Legacy pluggable transports for ``NodeRunner`` clusters (reference engine/transport/*)."""
from .base import Transport
from .local import LocalTransport
from .tcp import TcpTransport
from .tcp_simple import TcpMailbox, send_message

__all__ = ["Transport", "LocalTransport", "TcpTransport", "TcpMailbox", "send_message"]
