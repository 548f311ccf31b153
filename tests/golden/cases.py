"""Golden cases shared by make_golden.py (which runs the emulated reference) and the tests that replay them."""

# name -> scene kwargs, engine kwargs, number of frames
CASES = {
    # 8^3 blocks, colour, 5 % depth holes, short range so the chunk sphere covers the room
    "g8_color_holes": dict(
        scene=dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=10, spheres=((2.8, 1.5, 1.0, 0.4),), color=True, holes=0.05),
        vpb=8, vox_size=0.04, trunc=0.2, max_depth=3.0, frames=4),
    # 8^3 blocks, room straddling the origin (negative block coordinates), reference MaxDepth 10
    "g8_negative_coords": dict(
        scene=dict(width=200, height=160, room=(6.0, 5.0, 2.6), room_min=(-3.0, -2.5, -1.3), n_frames=12, spheres=((1.9, 0.4, -0.2, 0.5),)),
        vpb=8, vox_size=0.03, trunc=0.15, max_depth=10.0, frames=3),
    # the reference's own constants: 5^3 blocks, 3.6 cm voxels, 18 cm truncation (scene0220_02.yaml:50-52)
    "g5_reference_defaults": dict(
        scene=dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=10, spheres=((2.8, 1.5, 1.0, 0.4),), color=True, holes=0.02),
        vpb=5, vox_size=0.036, trunc=0.18, max_depth=4.0, frames=3),
}
