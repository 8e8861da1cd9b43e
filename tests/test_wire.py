"""N4 wire format (gaussian_renderer/network_gui.py:26-86): a fake SIBR viewer on a loopback socket exchanges one camera
request and one frame with gs_icp_slam_b200.wire.ViewerLink; framing and matrix flips checked against the reference's own
expressions."""
import json
import socket
import threading

import numpy as np
import torch


def _request(w, h):
    rng = np.random.default_rng(0)
    return {"resolution_x": w, "resolution_y": h, "train": 1, "fov_y": 0.9, "fov_x": 1.2, "z_near": 0.01, "z_far": 100.0,
            "shs_python": 0, "rot_scale_python": 1, "keep_alive": 1, "scaling_modifier": 1.0,
            "view_matrix": rng.standard_normal(16).tolist(), "view_projection_matrix": rng.standard_normal(16).tolist()}


def test_viewer_round_trip():
    from gs_icp_slam_b200.wire import ViewerLink, encode_frame

    link = ViewerLink("127.0.0.1", 0)  # any free port
    got = {}

    def viewer():
        s = socket.create_connection((link.host, link.port))
        for req in (_request(0, 0), _request(8, 6)):
            raw = json.dumps(req).encode("utf-8")
            s.sendall(len(raw).to_bytes(4, "little") + raw)
        def exact(n):
            buf = b""
            while len(buf) < n:
                buf += s.recv(n - len(buf))
            return buf

        frame = exact(8 * 6 * 3)
        n = int.from_bytes(exact(4), "little")
        got["frame"], got["verify"] = frame, exact(n).decode("ascii")
        s.close()

    t = threading.Thread(target=viewer)
    t.start()
    while not link.try_connect():
        pass
    assert link.receive() is None  # resolution 0 x 0: nothing to render
    cam = link.receive()
    req = _request(8, 6)
    # the reference's expressions (network_gui.py:73-77)
    v = torch.reshape(torch.tensor(req["view_matrix"]), (4, 4))
    v[:, 1] = -v[:, 1]
    v[:, 2] = -v[:, 2]
    p = torch.reshape(torch.tensor(req["view_projection_matrix"]), (4, 4))
    p[:, 1] = -p[:, 1]
    assert cam["width"] == 8 and cam["height"] == 6 and cam["do_training"] and not cam["do_shs_python"] and cam["do_rot_scale_python"]
    assert torch.equal(cam["world_view_transform"], v.float()) and torch.equal(cam["full_proj_transform"], p.float())
    img = torch.rand(3, 6, 8) * 1.4 - 0.2
    link.send(encode_frame(img), "synthetic/office0")
    t.join(timeout=10)
    link.close()
    want = (torch.clamp(img, 0, 1) * 255).byte().permute(1, 2, 0).contiguous().numpy().tobytes()
    assert got["frame"] == want and got["verify"] == "synthetic/office0"
