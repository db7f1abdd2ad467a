"""Wire-format shim of the NeRF service's scoring routes (nerfServer/core/views.py:94-139), framework free:
each handler takes the request body / query the Django view receives and returns the dict the view wraps in a
``JsonResponse``, so the unmodified planner client (plannerServer_*/core/interface2.py:46-84) can talk to this
engine through any HTTP front end.  The reference returns ``None`` (HTTP 500) for non-POST / non-GET methods;
here a wrong method raises ``ValueError``."""
import json

from .run_nerf import get_pose, get_poses  # noqa: F401


def get_uncertainty(controller, body, method="POST"):
    """views.py:94-113.  body: bytes/str JSON ``{"locations": [[x,y,z]..], "us": [..], "vs": [..]}`` (radians)
    -> ``{"uncers": [..]}``."""
    if method != "POST":
        raise ValueError("get_uncertainty expects POST")
    data = json.loads(body) if isinstance(body, (bytes, str)) else body
    locations, us, vs = data["locations"], data["us"], data["vs"]
    poses = get_poses(locations, us, vs)          # == [get_pose(locations[i], us[i], vs[i]) for i in ...] (views.py:104-107)
    uncers = controller.get_uncertainty(poses)
    return {"uncers": uncers.cpu().numpy().tolist()}


def get_surface_points(controller, query, method="GET"):
    """views.py:115-127.  query: mapping with x, y, z, radius, step -> ``{"points": [[x,y,z]..]}``."""
    if method != "GET":
        raise ValueError("get_surface_points expects GET")
    x, y, z = float(query["x"]), float(query["y"]), float(query["z"])
    points = controller.get_surface_points([x, y, z], float(query["radius"]), float(query["step"]))
    return {"points": points.cpu().numpy().tolist()}


def get_ray_depth(controller, body, method="POST"):
    """views.py:129-139.  body JSON ``{"locations": [...], "directions": [...]}`` -> ``{"depths": [...]}``."""
    if method != "POST":
        raise ValueError("get_ray_depth expects POST")
    data = json.loads(body) if isinstance(body, (bytes, str)) else body
    depths = controller.get_rays_depth(data["locations"], data["directions"])
    return {"depths": depths.cpu().numpy().tolist()}


def planner_client(controller):
    """The planner side of the wire (plannerServer_*/core/interface2.py:46-60) without the HTTP hop:
    ``get_uncertainty(locations, us, vs) -> (uncers, [], [])`` built on ``controller``, encoding and decoding the
    same JSON bodies the two services exchange.  This is the callable ``planner.sample_direction`` expects."""
    def client(locations, us, vs):
        body = json.dumps({"locations": [list(map(float, x)) for x in locations], "us": [float(u) for u in us],
                           "vs": [float(v) for v in vs]})
        reply = json.loads(json.dumps(get_uncertainty(controller, body)))
        return reply["uncers"], [], []
    return client
