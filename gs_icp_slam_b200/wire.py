"""SURVEY.md §8f row N4 — the SIBR remote-viewer wire format of gaussian_renderer/network_gui.py:26-86, as a small class
instead of module globals (several links per process, no socket created at import time).  Host code only: the frame that goes
out is whatever the rasterizer produced.

Protocol (little endian):
  viewer -> us   u32 length, then `length` bytes of UTF-8 JSON: resolution_x/y, train, fov_y, fov_x, z_near, z_far, shs_python,
                 rot_scale_python, keep_alive, scaling_modifier, view_matrix[16], view_projection_matrix[16]
  us -> viewer   optional raw frame (H*W*3 bytes, RGB8, row major), then u32 length + that many ASCII bytes (the "verify" string,
                 the dataset path in the reference)
The camera matrices arrive in the viewer's y-up / z-back convention: columns 1 and 2 of the view matrix and column 1 of the
view-projection matrix are negated (network_gui.py:73-77)."""
import json
import socket

import torch


def encode_frame(image_chw):
    """[3, H, W] float image in [0, 1] (any device) -> H*W*3 RGB8 bytes, the payload the viewer expects
    (train.py of the 3DGS code base: clamp, *255, byte, HWC)."""
    img = (torch.clamp(image_chw.detach(), min=0.0, max=1.0) * 255).to(torch.uint8).permute(1, 2, 0).contiguous()
    return img.cpu().numpy().tobytes()


def decode_camera(message, device="cpu"):
    """The JSON request as a dict of plain values / tensors, or None for the viewer's "no frame wanted" message
    (resolution 0 x 0).  Same fields and flips as network_gui.receive (:54-84); MiniCam construction is left to the caller."""
    width, height = int(message["resolution_x"]), int(message["resolution_y"])
    if width == 0 or height == 0:
        return None
    view = torch.tensor(message["view_matrix"], dtype=torch.float32).reshape(4, 4).to(device)
    view[:, 1] = -view[:, 1]
    view[:, 2] = -view[:, 2]
    proj = torch.tensor(message["view_projection_matrix"], dtype=torch.float32).reshape(4, 4).to(device)
    proj[:, 1] = -proj[:, 1]
    return {"width": width, "height": height, "fovy": message["fov_y"], "fovx": message["fov_x"], "znear": message["z_near"],
            "zfar": message["z_far"], "do_training": bool(message["train"]), "do_shs_python": bool(message["shs_python"]),
            "do_rot_scale_python": bool(message["rot_scale_python"]), "keep_alive": bool(message["keep_alive"]),
            "scaling_modifier": message["scaling_modifier"], "world_view_transform": view, "full_proj_transform": proj}


class ViewerLink:
    """Listener + one viewer connection (network_gui.init / try_connect / read / send / receive)."""

    def __init__(self, host="127.0.0.1", port=6009):
        self.listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.listener.bind((host, port))
        self.listener.listen()
        self.listener.settimeout(0)
        self.host, self.port = self.listener.getsockname()
        self.conn, self.addr = None, None

    def try_connect(self):
        """Non-blocking accept, like the reference; True once a viewer is attached."""
        if self.conn is None:
            try:
                self.conn, self.addr = self.listener.accept()
                self.conn.settimeout(None)
            except (BlockingIOError, socket.timeout, OSError):
                pass
        return self.conn is not None

    def _recv_exact(self, n):
        buf = bytearray()
        while len(buf) < n:  # the reference calls recv(n) once and may get a short read on large messages
            chunk = self.conn.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("viewer closed the connection")
            buf.extend(chunk)
        return bytes(buf)

    def read(self):
        length = int.from_bytes(self._recv_exact(4), "little")
        return json.loads(self._recv_exact(length).decode("utf-8"))

    def send(self, frame_bytes, verify):
        if frame_bytes is not None:
            self.conn.sendall(frame_bytes)
        self.conn.sendall(len(verify).to_bytes(4, "little"))
        self.conn.sendall(bytes(verify, "ascii"))

    def receive(self, device="cpu"):
        return decode_camera(self.read(), device)

    def close(self):
        for s in (self.conn, self.listener):
            if s is not None:
                try:
                    s.close()
                except OSError:
                    pass
        self.conn = None
