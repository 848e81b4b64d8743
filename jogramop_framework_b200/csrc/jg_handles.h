// jg_handles.h -- the host-side robot / scene handles of the C-ABI (shared by jg_abi.cu and jg_ingest.cpp).
#pragma once
#include <string>
#include <vector>

struct jg_robot {
    int L = 0, S = 0, dof = 0, K = 0;
    std::vector<int> joint_type, joint_parent, joint_qidx, shape_link, shape_off, plane_off, self_pairs;
    std::vector<double> joint_origin, joint_axis, q_limits, shape_margin, core, plane_v, plane_margin;
    // filled by jg_robot_load (empty for robots made with jg_robot_create)
    std::vector<std::string> link_names, joint_names;
    int ee_link = -1;
};

struct jg_scene {
    int P = 0;
    std::vector<int> off;
    std::vector<double> v, margin;
    bool is_mesh = false;              // triangle soup behind an LBVH instead of convex pieces
    std::vector<float> mesh_v;         // [n,3]
    std::vector<int> mesh_t;           // [m,3]
    double mesh_margin = 0.001;
    bool has_plane = false;
    double plane[4] = {0, 0, 1, 0};
    // filled by jg_scene_load: body index of every piece (objects first, then bg objects: create_cpp_export.py:41)
    std::vector<int> piece_body;
    std::vector<std::string> body_names;
    int n_target_bodies = 0;
};
