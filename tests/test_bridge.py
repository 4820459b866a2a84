"""CPU: the coordinator wire format (reference include/mpl/packet.hpp, buffer.hpp: big-endian, type:u32 size:u32
payload) and the link that speaks it, against a stand-in coordinator on a loopback socket."""
import socket
import struct
import threading

import numpy as np
import pytest

from mplambda_b200 import bridge as B


def test_packet_bytes_are_the_reference_layout():
    # Hello / Done: packet.hpp:127-135,158-166 -- 16 bytes, id last
    assert B.encode_hello(0x0102030405060708) == bytes.fromhex("3864caca" "00000010" "0102030405060708")
    assert B.encode_done(7) == bytes.fromhex("6672e31a" "00000010" "0000000000000007")
    # Path<SE3 double>: TYPE = PATH_SE3 + sizeof(double)/8 (packet.hpp:188), cost, millis, then per state the
    # quaternion coefficients x y z w (buffer.hpp:182-184: q.coeffs()) and the translation
    path = np.array([[0.0, 0.0, 0.0, 1.0, 1.5, -2.0, 3.0], [0.5, 0.5, 0.5, 0.5, 0.0, 0.25, -0.125]])
    pkt = B.encode_path(12.5, 250, path)
    want = struct.pack(">II", 0xA9DB6E7D + 1, 8 + 8 + 4 + 2 * 56) + struct.pack(">dI", 12.5, 250) + struct.pack(">14d", *path.reshape(-1))
    assert pkt == want
    # Path<Matrix<double, 8, 1>>: TYPE = PATH_RVD + dim (packet.hpp:176)
    q = np.arange(16, dtype=np.float64).reshape(2, 8) / 8
    pkt = B.encode_path(1.0, 3, q)
    assert pkt[:8] == struct.pack(">II", 0xB11B0C45 + 0x100 + 8, 8 + 12 + 128)
    assert pkt[20:] == struct.pack(">16d", *q.reshape(-1))


def test_parse_round_trip_partial_and_errors():
    path = np.random.default_rng(0).normal(size=(5, 7))
    stream = B.encode_hello(42) + B.encode_path(3.25, 17, path) + B.encode_done(42)
    # byte by byte: nothing is returned until a packet is complete (packet.hpp:257-275)
    buf, got = bytearray(), []
    for b in stream:
        buf.append(b)
        pkt, used = B.parse(buf)
        if pkt is not None:
            got.append(pkt)
            del buf[:used]
    assert [g[0] for g in got] == ["hello", "path", "done"] and got[0][1] == 42 and got[2][1] == 42
    assert got[1][1] == 3.25 and got[1][2] == 17 and (got[1][3] == path).all() and got[1][4] == 8
    # the float variants the reference also parses (packet.hpp:304-315)
    f32 = struct.pack(">II", 0xA9DB6E7D, 8 + 8 + 28) + struct.pack(">fI", 2.0, 9) + struct.pack(">7f", *range(7))
    pkt, used = B.parse(f32)
    assert used == len(f32) and pkt[1] == 2.0 and pkt[4] == 4 and (pkt[3] == np.arange(7.0)).all()
    rvf = struct.pack(">II", 0xB11B0C45 + 8, 8 + 8 + 32) + struct.pack(">fI", 1.0, 1) + struct.pack(">8f", *range(8))
    assert B.parse(rvf)[0][3].shape == (1, 8)
    with pytest.raises(B.ProtocolError):   # bad packet type (packet.hpp:316-317)
        B.parse(struct.pack(">IIQ", 0x12345678, 16, 0))
    with pytest.raises(B.ProtocolError):   # maximum packet size exceeded (packet.hpp:265-266)
        B.parse(struct.pack(">II", B.HELLO, B.MAX_PACKET_SIZE + 1))
    with pytest.raises(B.ProtocolError):   # invalid path packet size (packet.hpp:216-217)
        B.parse(struct.pack(">II", 0xA9DB6E7D + 1, 8 + 12 + 55) + bytes(12 + 55))
    with pytest.raises(B.ProtocolError):
        B.encode_path(1.0, 0, np.zeros((20000, 7)))


def _coordinator(server, log, relay):
    """What mpl_coordinator does for one group (mpl_coordinator.cpp:529-589): reads the lambda's packets, relays a
    foreign path after the first one it receives, answers Done with Done."""
    conn, _ = server.accept()
    buf = bytearray()
    sent_foreign = False
    while True:
        chunk = conn.recv(65536)
        if not chunk:
            break
        buf += chunk
        while True:
            pkt, used = B.parse(buf)
            if pkt is None:
                break
            del buf[:used]
            log.append(pkt)
            if pkt[0] == "path" and not sent_foreign:
                conn.sendall(B.encode_path(relay[0], 5, relay[1]))
                sent_foreign = True
            if pkt[0] == "done":
                conn.sendall(B.encode_done(pkt[1]))
                conn.close()
                return


def test_link_against_a_stand_in_coordinator():
    server = socket.socket()
    server.bind(("127.0.0.1", 0))
    server.listen(1)
    log = []
    foreign = (1.5, np.array([[0, 0, 0, 1.0, 1, 2, 3], [0, 0, 0, 1.0, 4, 5, 6]]))
    t = threading.Thread(target=_coordinator, args=(server, log, foreign), daemon=True)
    t.start()
    link = B.CoordinatorLink(0xABCDEF, "127.0.0.1", server.getsockname()[1])
    mine = np.array([[0, 0, 0, 1.0, 0, 0, 0], [0, 0, 0, 1.0, 9, 9, 9]])
    link.send_path(4.0, 120, mine)
    got = []
    for _ in range(50):
        got += link.poll(timeout=0.1)
        if got:
            break
    assert got and got[0][0] == "path" and got[0][1] == 1.5 and (got[0][3] == foreign[1]).all()
    link.send_done()
    for _ in range(50):
        got += link.poll(timeout=0.1)
        if any(g[0] == "done" for g in got):
            break
    t.join(5)
    link.close()
    server.close()
    assert link.done and [p[0] for p in log] == ["hello", "path", "done"]
    assert log[0][1] == 0xABCDEF and log[1][1] == 4.0 and log[1][2] == 120 and (log[1][3] == mine).all()
