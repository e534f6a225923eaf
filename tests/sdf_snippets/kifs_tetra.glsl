// kaleidoscopic IFS tetrahedron: conditional swizzle swaps with negation
float sdf(in vec3 p) {
    vec3 z = p;
    const float scale = 2.0;
    const vec3 offset = vec3(0.7);
    int n = 0;
    for (int i = 0; i < 7; i++) {
        if (z.x + z.y < 0.0) z.xy = -z.yx;
        if (z.x + z.z < 0.0) z.xz = -z.zx;
        if (z.y + z.z < 0.0) z.zy = -z.yz;
        z = z * scale - offset * (scale - 1.0);
        n++;
    }
    return length(z) * pow(scale, -float(n)) - 0.01;
}

float sdfmaterial(in vec3 p) {
    return 0.0;
}
