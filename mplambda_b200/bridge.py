"""TCP bridge from the planner shards to the reference's coordinator.

The reference's lambdas talk to `mpl_coordinator` over one TCP connection each (reference src/mpl/comm.cpp): on
connect they send `Hello(problemId)` (comm.cpp:31), every improved solution goes out as a `Path` packet
(include/mpl/comm.hpp:80-106), `Done(problemId)` ends the job (comm.cpp:144-152); the coordinator relays paths to the
initiator -- and, for C-FOREST, to the other lambdas of the group -- and broadcasts Done
(src/mpl_coordinator.cpp:529-589).  Between the GPUs of one node that exchange is a collective
(sharding.SolutionExchange); this module is the other end: the winning rank speaks the coordinator's wire format, so a
multi-GPU job looks like one (very fast) lambda to an unchanged coordinator and its client.

Wire format (include/mpl/packet.hpp, include/mpl/buffer.hpp:97-116,164-203): every number big-endian; a packet is
`type:u32 size:u32 payload`, size counting the 8 header bytes.
  Hello 0x3864caca / Done 0x6672e31a : id:u64                                          (packet.hpp:104-168)
  Path  0xa9db6e7d + sizeof(S)/8     : cost:S millis:u32 states (qx qy qz qw tx ty tz)  (SE3, packet.hpp:185-194)
  Path  0xb11b0c45 [+0x100] + dim    : cost:S millis:u32 states (dim scalars)           (RVF / RVD, packet.hpp:173-183)
Only the float64 variants are sent (the drop-in scenarios are Scenario<double>); all four are parsed.
"""
import socket
import struct

import numpy as np

PROBLEM, HELLO, DONE = 0x8179E3F1, 0x3864CACA, 0x6672E31A
PATH_SE3, PATH_RVF = 0xA9DB6E7D, 0xB11B0C45
PATH_RVD = PATH_RVF + 0x100
MAX_PACKET_SIZE = 1024 * 1024          # packet.hpp:21
DEFAULT_PORT = 0x415E                  # comm.hpp:13


class ProtocolError(RuntimeError):
    """packet.hpp:26-32 protocol_error"""


def encode_hello(problem_id):
    return struct.pack(">IIQ", HELLO, 16, int(problem_id))


def encode_done(problem_id):
    return struct.pack(">IIQ", DONE, 16, int(problem_id))


def encode_path(cost, solve_millis, path):
    """Path<State> for State = SE3 tuple (path [n, 7]) or Eigen::Matrix<double, dim, 1> (path [n, dim]), float64."""
    p = np.ascontiguousarray(path, dtype=np.float64)
    if p.ndim != 2:
        raise ValueError("path must be [n, dim]")
    dim = p.shape[1]
    ptype = PATH_SE3 + 1 if dim == 7 else PATH_RVD + dim
    body = struct.pack(">dI", float(cost), int(solve_millis) & 0xFFFFFFFF) + p.astype(">f8").tobytes()
    size = 8 + len(body)
    if size > MAX_PACKET_SIZE:
        raise ProtocolError("maximum packet size exceeded: %d" % size)
    return struct.pack(">II", ptype, size) + body


def parse(buf):
    """packet.hpp:254-325: -> (packet or None, bytes consumed).  A packet is ("hello" | "done", id) or
    ("path", cost, millis, states [n, dim] float64, scalar_bytes); None: more bytes are needed."""
    if len(buf) < 8:
        return None, 0
    ptype, size = struct.unpack_from(">II", buf, 0)
    if size > MAX_PACKET_SIZE:
        raise ProtocolError("maximum packet size exceeded: %d" % size)
    if size < 8:
        raise ProtocolError("invalid packet size: %d" % size)
    if len(buf) < size:
        return None, 0
    body = bytes(buf[8:size])
    if ptype in (HELLO, DONE):
        if len(body) < 8:
            raise ProtocolError("short %s packet" % ("hello" if ptype == HELLO else "done"))
        return ("hello" if ptype == HELLO else "done", struct.unpack_from(">Q", body)[0]), size
    if ptype in (PATH_SE3, PATH_SE3 + 1):
        scalar, dim = (4, 7) if ptype == PATH_SE3 else (8, 7)
    elif ptype == PATH_RVF + 8:
        scalar, dim = 4, 8
    elif ptype == PATH_RVD + 8:
        scalar, dim = 8, 8
    else:
        raise ProtocolError("bad packet type: %d" % ptype)
    head = scalar + 4
    if len(body) < head or (len(body) - head) % (scalar * dim) != 0:
        raise ProtocolError("invalid path packet size: %d" % (len(body) - head))
    cost = struct.unpack_from(">f" if scalar == 4 else ">d", body)[0]
    millis = struct.unpack_from(">I", body, scalar)[0]
    states = np.frombuffer(body, dtype=">f4" if scalar == 4 else ">f8", offset=head).astype(np.float64).reshape(-1, dim)
    return ("path", cost, millis, states, scalar), size


class CoordinatorLink:
    """One lambda's connection (mpl::Comm): connect -> Hello; send_path / send_done; poll() hands back what the
    coordinator relays (paths of the other group members, Done)."""

    def __init__(self, problem_id, host=None, port=DEFAULT_PORT, sock=None, timeout=5.0):
        self.problem_id = int(problem_id)
        self.done = False
        self._rbuf = bytearray()
        self._sock = sock if sock is not None else socket.create_connection((host, port), timeout=timeout)
        self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._sock.sendall(encode_hello(self.problem_id))           # comm.cpp:31

    def send_path(self, cost, solve_millis, path):
        self._sock.sendall(encode_path(cost, solve_millis, path))    # comm.hpp:80-106

    def send_done(self):
        if not self.done:
            self._sock.sendall(encode_done(self.problem_id))         # comm.cpp:144-152
            self.done = True

    def poll(self, timeout=0.0):
        """Packets that have arrived (non-blocking unless `timeout` > 0); a Done sets `done` (comm.cpp:139-142)."""
        out = []
        self._sock.settimeout(timeout if timeout > 0 else 0.0)
        try:
            while True:
                chunk = self._sock.recv(65536)
                if not chunk:
                    self.done = True
                    break
                self._rbuf += chunk
                if timeout > 0:
                    self._sock.settimeout(0.0)
        except (BlockingIOError, socket.timeout, TimeoutError):
            pass
        while True:
            pkt, used = parse(self._rbuf)
            if pkt is None:
                break
            del self._rbuf[:used]
            if pkt[0] == "done":
                self.done = True
            out.append(pkt)
        return out

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass


class BridgedExchange:
    """SolutionExchange + the coordinator link on rank 0 of the group: what the ranks agree on is forwarded upstream
    (each improvement once), what the coordinator relays from other lambdas joins the next exchange as rank 0's
    candidate, and a Done from upstream stops every rank (mpl_coordinator.cpp:538-589)."""

    def __init__(self, exchange, link=None):
        self.ex, self.link = exchange, link          # link: CoordinatorLink on the group's rank 0, None elsewhere
        self.best_cost = float("inf")
        self._foreign = (float("inf"), None)

    def publish(self, cost, path, solve_millis=0):
        if self.link is not None:
            for pkt in self.link.poll():
                if pkt[0] == "path" and pkt[1] < self._foreign[0]:
                    self._foreign = (pkt[1], pkt[3])
            if self._foreign[0] < cost:
                cost, path = self._foreign
        best_cost, best_rank, best_path = self.ex.publish(cost, path)
        if self.link is not None and best_path is not None and best_cost < self.best_cost and best_cost < self._foreign[0]:
            self.link.send_path(best_cost, solve_millis, best_path)
        self.best_cost = min(self.best_cost, best_cost)
        return best_cost, best_rank, best_path

    def any_done(self, done):
        upstream = self.link is not None and (bool(self.link.poll()) or True) and self.link.done
        stop = self.ex.any_done(bool(done) or upstream)
        if stop and self.link is not None:
            self.link.send_done()
        return stop
