// gyroid shell clipped to a sphere: vector sin / cos and a three-letter swizzle of an expression
float sdf(in vec3 p) {
    float s = 9.0;
    float g = dot(sin(p * s), cos((p * s).zxy)) / s;
    float shell = abs(g) - 0.03;
    return max(shell * 0.7, length(p) - 0.9);
}

float sdfmaterial(in vec3 p) {
    return 0.0;
}
