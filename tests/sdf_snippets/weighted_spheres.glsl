// constant arrays through array constructors, indexed in a loop
float sdf(in vec3 p) {
    const float radius[4] = float[4](0.35, 0.25, 0.2, 0.15);
    const vec3 centre[4] = vec3[4](vec3(0.0, 0.0, 0.0), vec3(0.45, 0.1, 0.0), vec3(-0.3, 0.35, 0.2), vec3(0.0, -0.4, -0.3));
    float d = MAXDIST;
    for (int i = 0; i < 4; i++) {
        d = smin(d, length(p - centre[i]) - radius[i]);
    }
    return d;
}

float sdfmaterial(in vec3 p) {
    float w[] = float[](0.0, 1.0, 2.0);
    return w[int(clamp(floor(p.y * 2.0 + 1.5), 0.0, 2.0))];
}
