// Position space of one stride-1 convolution (see conv_shift.cu): the frame of its padded / upsampled input and the
// storage geometry of the planar operand planes [plane][8-channel group][position][8] built on it.
#pragma once
struct PosFrame {
    int N, Hp, Wp;        // frame of padded positions, q = (n*Hp + yp)*Wp + xp
    int G;                // 8-channel groups stored per plane (even)
    int lead;             // zero positions stored in front of q = 0
    long long QA;         // positions stored per (plane, group)
};
