def euler_to_vec(yaw, pitch):
    raise NotImplementedError("GUI helper, not on the hot path")


def vec_to_euler(v):
    raise NotImplementedError("GUI helper, not on the hot path")
